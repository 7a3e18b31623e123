/*
 * pcb_kernels.cu -- kernel launchers of libpolycap_b200 (sm_100a).
 *
 *   pcb_trace_kernel     persistent warp-synchronous tracer            (pcb_trace.cuh)
 *   wave_scan_*          exclusive scan of the per-id exit flags: exit photons get their
 *                        record slot in id order (deterministic output, reference: slot j of
 *                        source.c:745)
 *   wave_tally_*         fixed-order reduction of exit weights and counters for ids below
 *                        the cut (source.c:893-896, 963-980)
 *   wave_compact_*       dense per-id records -> packed struct _polycap_images arrays
 *   dfma_peak_kernel     FP64 FMA throughput microbenchmark (roofline denominator)
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "pcb_gpu.h"
#include "pcb_leak.h"
#include "pcb_trace.cuh"
#include "pcb_trace2.cuh"

extern "C" void pcb_energy_constants(size_t n_energies, const double *energies, const double *amu, const double *scatf, double density, double *out);

#define COSPI_6_HOST 0.86602540378443864676

static thread_local std::string g_last_error;

#define CK(call)                                                                                              \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                      \
		if (e_ != cudaSuccess) {                                                                      \
			char b_[512];                                                                         \
			snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
			g_last_error = b_;                                                                    \
			return -1;                                                                            \
		}                                                                                             \
	} while (0)

extern "C" const char *pcb_gpu_last_error(void) { return g_last_error.c_str(); }

extern "C" int pcb_gpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		g_last_error = "cudaGetDeviceCount failed: no usable CUDA device/driver";
		return 0;
	}
	return n;
}

/* ------------------------------------------------------------------ device optic */

struct pcb_gpu_optic {
	int device = 0;
	int sm_count = 0;
	int nmax = 0, nE = 0;
	PcbOptic o{};
	double *d_tables = nullptr; /* z, cap, ext, hexz, rm: 5 x tb bytes */
	double *d_leak_tables = nullptr; /* leak transport only: per z segment d ext/dz, d cap/dz, step length cap/10 and its reciprocal: 4 x tb bytes */
	double *d_ec = nullptr;
	double *d_wscratch = nullptr;
	double *d_hist = nullptr; /* deferred weights: reflection history of the pool slots */
	int defer_min = 48;	  /* deferred weights for grids of at least this many energies (source-driven runs, cheap weights) */
	int hk = 0;		  /* history entries per flush (0: from the energy count) */
	PcbWorkCounters *d_wc = nullptr;
	size_t table_bytes = 0; /* tb */
	int threads = PCB_LB_THREADS, ctas_per_sm = PCB_LB_CTAS;
	int thr_cand = 12, thr_refl = 12, thr_new = 8;
	int tables_in_smem = 1;
	int exact_weights = 0;
	int kernel_version = 2; /* 2: CTA task-queue tracer (pcb_trace2.cuh, default), 1: warp-voted tracer (pcb_trace.cuh) */
	int monotonic = 1; /* z strictly increasing */
	/* pinned host memory: the staged tables/constants (re-uploadable) and a grow-only staging area for results */
	double *h_tables = nullptr, *h_ec = nullptr;
	size_t ec_bytes = 0;
	void *h_stage = nullptr;
	size_t h_stage_bytes = 0;
	void *d_result = nullptr; /* exit records + tallies of one polycap_source_get_transmission_efficiencies call (grow-only) */
	size_t d_result_bytes = 0;
	uint64_t hint_key = 0; /* exit fraction seen by the last call with this source geometry: sizes the first wave of the next */
	double hint_frac = 0;
	int64_t hint_exits = 0;
	double cap0 = 0, ext0 = 0;
	/* wave buffers (grown on demand) */
	int64_t wave_cap = 0;
	int8_t *d_status = nullptr;
	double *d_fields = nullptr;
	int *d_irefl = nullptr;
	double *d_weights = nullptr;
	int64_t *d_rank = nullptr;
	int64_t *d_blocksum = nullptr;
	int64_t *d_cutinfo = nullptr; /* [0] cut, [1] exits below cut, [2] total exits */
	double *d_partial = nullptr;
	int64_t *d_cpartial = nullptr;
	int64_t last_wave_n = 0;
	/* replay buffers */
	int64_t replay_cap = 0;
	double *d_rin = nullptr;
	/* two-phase waves (multi-device runs, pipelined record download): the optic's own streams and a pinned word block the
	 * device copies wave counts into ([0] ids below the cut, [1] exits below the cut, [2] exits of the whole wave) */
	double *d_reduce = nullptr; /* output of tally_reduce_peers (nE doubles + the counter block) */
	cudaStream_t s_main = nullptr, s_copy = nullptr;
	cudaEvent_t ev_fin = nullptr, ev_copy = nullptr;
	int64_t *h_pin = nullptr;
	/* instrumentation: kernels launched, and (when timing is on) CUDA-event pairs around every trace kernel */
	int64_t n_kernels = 0, n_trace = 0;
	int timing = 0;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending;
	std::vector<cudaEvent_t> ev_pool;
	double trace_ms = 0;
	int64_t trace_timed = 0;
};

static int pick_device(int device)
{
	if (device >= 0)
		return device;
	const char *env = getenv("POLYCAP_B200_DEVICE");
	if (env && *env)
		return atoi(env);
	int cur = 0;
	if (cudaGetDevice(&cur) != cudaSuccess)
		return 0;
	return cur;
}

extern "C" int pcb_gpu_resolve_device(int device) { return pick_device(device); }

extern "C" pcb_gpu_optic *pcb_gpu_optic_create(int device, const pcb_optic_desc *d)
{
	int ndev = pcb_gpu_device_count();
	if (ndev <= 0)
		return nullptr;
	device = pick_device(device);
	if (device >= ndev) {
		g_last_error = "requested CUDA device does not exist";
		return nullptr;
	}
	if (cudaSetDevice(device) != cudaSuccess) {
		g_last_error = "cudaSetDevice failed";
		return nullptr;
	}
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
		g_last_error = "cudaGetDeviceProperties failed";
		return nullptr;
	}
	if (prop.major != 10 || prop.minor != 0) { /* the library carries sm_100a code only (no PTX) */
		g_last_error = "polycap-b200 needs an sm_100a (Blackwell B200) device";
		return nullptr;
	}
	pcb_gpu_optic *g = new pcb_gpu_optic();
	g->device = device;
	g->sm_count = prop.multiProcessorCount;
	g->nmax = d->nmax;
	g->nE = (int)d->n_energies;
	const int n = d->nmax + 1;
	const size_t tb = ((size_t)(n + 1) * 8 + 15) & ~(size_t)15; /* one padded entry: the reference can read index nmax+1 */
	g->table_bytes = tb;
	const size_t tn = tb / 8;

	double n_shells = round(sqrt(12. * d->n_cap - 3.) / 6. - 0.5);
	std::vector<double> h(5 * tn, 0.0);
	double *hz = h.data(), *hcap = hz + tn, *hext = hcap + tn, *hhex = hext + tn, *hrm = hhex + tn;
	const double eps = 1e-9;
	double m_safe = 1e300, m_deep = 1e300;
	bool mono_ok = true;
	for (int i = 0; i < n; i++) {
		hz[i] = d->z[i];
		hcap[i] = d->cap[i];
		hext[i] = d->ext[i];
		hhex[i] = d->ext[i] / (2. * COSPI_6_HOST * (n_shells + 1));
		hrm[i] = d->cap[i] * d->cap[i] * (1.0 - eps);
		double apo = sqrt((d->ext[i] * d->ext[i]) - ((d->ext[i] / 2.) * (d->ext[i] / 2.)));
		/* interior capillary: bore wall plus the largest lateral excursion of a back-extrapolated restart point
		 * (one segment length at tan(theta) <= 0.01, see trace_begin) stays inside the outer hexagon at every z */
		double dzmax = 0;
		if (i > 0) dzmax = fmax(dzmax, d->z[i] - d->z[i - 1]);
		if (i + 1 < n) dzmax = fmax(dzmax, d->z[i + 1] - d->z[i]);
		double room = (apo * (1.0 - eps) - d->cap[i] - 0.01 * dzmax) / hhex[i];
		double room_deep = (apo * (1.0 - eps) - d->cap[i] - 0.25 * dzmax) / hhex[i];
		if (!(room_deep < 1e300))
			room_deep = -1;
		if (room_deep < m_deep)
			m_deep = room_deep;
		if (!(room < 1e300))
			room = -1;
		if (room < m_safe)
			m_safe = room;
		if (!(d->cap[i] + 0.01 * dzmax < apo * (1.0 - eps) && d->cap[i] < d->ext[i] * (1.0 - eps) && d->cap[i] * d->cap[i] < d->ext[i] * (1.0 - eps)))
			mono_ok = false;
		if (!(d->cap[i] > 0) || !(d->ext[i] > 0))
			m_safe = -1;
	}
	for (size_t i = n; i < tn; i++) {
		hz[i] = hz[n - 1]; hcap[i] = hcap[n - 1]; hext[i] = hext[n - 1]; hhex[i] = hhex[n - 1]; hrm[i] = 0.0;
	}
	for (int i = 0; i + 1 < n; i++)
		if (!(d->z[i + 1] > d->z[i]))
			m_safe = -1; /* non-monotonic z: only the exact path is trusted */
	for (int i = 0; i + 1 < n; i++)
		if (!(d->z[i + 1] > d->z[i]))
			g->monotonic = 0;

	PcbOptic &o = g->o;
	o.nmax = d->nmax;
	o.tstride = (int)tn;
	o.n_energies = g->nE;
	o.mono = (n_shells == 0.);
	o.all_slow = getenv("POLYCAP_B200_EXACT_CHECKS") ? 1 : 0;
	o.n_shells = n_shells;
	o.sig_rough = d->sig_rough;
	o.hex_den = 2. * COSPI_6_HOST * (n_shells + 1);
	o.m_safe = o.mono ? (mono_ok && m_safe > 0 ? 1.0 : -1.0) : m_safe * (1.0 - eps);
	o.m_deep = (m_safe > 0 && !o.mono) ? m_deep * (1.0 - eps) : -1.0;

	std::vector<double> ec((size_t)EC_ROWS * g->nE);
	pcb_energy_constants(d->n_energies, d->energies, d->amu, d->scatf, d->density, ec.data());

	bool ok = cudaMalloc(&g->d_tables, 5 * tb) == cudaSuccess && cudaMalloc(&g->d_ec, ec.size() * 8) == cudaSuccess &&
		  cudaMalloc(&g->d_wc, sizeof(PcbWorkCounters)) == cudaSuccess;
	g->ec_bytes = ec.size() * 8;
	if (ok)
		ok = cudaHostAlloc((void **)&g->h_tables, 5 * tb, cudaHostAllocDefault) == cudaSuccess &&
		     cudaHostAlloc((void **)&g->h_ec, g->ec_bytes, cudaHostAllocDefault) == cudaSuccess;
	if (ok) {
		memcpy(g->h_tables, h.data(), 5 * tb);
		memcpy(g->h_ec, ec.data(), g->ec_bytes);
		ok = cudaMemcpy(g->d_tables, g->h_tables, 5 * tb, cudaMemcpyHostToDevice) == cudaSuccess &&
		     cudaMemcpy(g->d_ec, g->h_ec, g->ec_bytes, cudaMemcpyHostToDevice) == cudaSuccess;
	}
	if (ok) { /* tables of the wall march (pcb_leak_core.cuh::LkCtx), host IEEE division = the bits the march's own expressions give */
		std::vector<double> lt(4 * tn, 0.0);
		double *lse = lt.data(), *lsc = lse + tn, *lsl = lsc + tn, *lis = lsl + tn;
		for (int i = 0; i < n; i++) {
			if (i + 1 < n) {
				lse[i] = (d->ext[i + 1] - d->ext[i]) / (d->z[i + 1] - d->z[i]);
				lsc[i] = (d->cap[i + 1] - d->cap[i]) / (d->z[i + 1] - d->z[i]);
			}
			lsl[i] = d->cap[i] / 10.;
			lis[i] = 1.0 / lsl[i];
		}
		for (size_t i = n; i < tn; i++) {
			lsl[i] = lsl[n - 1];
			lis[i] = lis[n - 1];
		}
		ok = cudaMalloc(&g->d_leak_tables, 4 * tb) == cudaSuccess && cudaMemcpy(g->d_leak_tables, lt.data(), 4 * tb, cudaMemcpyHostToDevice) == cudaSuccess;
	}
	if (!ok) {
		g_last_error = std::string("device allocation/upload failed: ") + cudaGetErrorString(cudaGetLastError());
		pcb_gpu_optic_destroy(g);
		return nullptr;
	}
	o.z = g->d_tables;
	o.cap = g->d_tables + tn;
	o.ext = g->d_tables + 2 * tn;
	o.hexz = g->d_tables + 3 * tn;
	o.rm = g->d_tables + 4 * tn;
	o.ec = g->d_ec;
	g->cap0 = d->cap[0];
	g->ext0 = d->ext[0];
	g->tables_in_smem = (5 * tb <= 160 * 1024) ? 1 : 0;
	const char *e;
	if ((e = getenv("POLYCAP_B200_THR_CAND"))) g->thr_cand = atoi(e);
	if ((e = getenv("POLYCAP_B200_THR_REFL"))) g->thr_refl = atoi(e);
	if ((e = getenv("POLYCAP_B200_THR_NEW"))) g->thr_new = atoi(e);
	if ((e = getenv("POLYCAP_B200_CTAS_PER_SM"))) g->ctas_per_sm = atoi(e);
	if ((e = getenv("POLYCAP_B200_EXACT_FRESNEL"))) g->exact_weights = atoi(e);
	if ((e = getenv("POLYCAP_B200_KERNEL"))) g->kernel_version = atoi(e);
	if ((e = getenv("POLYCAP_B200_DEFER_MIN"))) g->defer_min = atoi(e); /* 0: never defer */
	if ((e = getenv("POLYCAP_B200_DEFER_HK"))) g->hk = atoi(e);
	if (t2_smem_bytes((int)tn) > 227 * 1024)
		g->kernel_version = 1; /* profile too long for the shared-memory pool */
	return g;
}

extern "C" void pcb_gpu_optic_destroy(pcb_gpu_optic *g)
{
	if (!g)
		return;
	cudaSetDevice(g->device);
	cudaFree(g->d_tables); cudaFree(g->d_leak_tables); cudaFree(g->d_ec); cudaFree(g->d_wscratch); cudaFree(g->d_hist); cudaFree(g->d_wc);
	cudaFree(g->d_status); cudaFree(g->d_fields); cudaFree(g->d_irefl); cudaFree(g->d_weights);
	cudaFree(g->d_rank); cudaFree(g->d_blocksum); cudaFree(g->d_cutinfo); cudaFree(g->d_partial); cudaFree(g->d_cpartial);
	cudaFree(g->d_rin);
	cudaFreeHost(g->h_tables);
	cudaFreeHost(g->h_ec);
	cudaFreeHost(g->h_stage);
	cudaFreeHost(g->h_pin);
	cudaFree(g->d_reduce);
	if (g->s_main) cudaStreamDestroy(g->s_main);
	if (g->s_copy) cudaStreamDestroy(g->s_copy);
	if (g->ev_fin) cudaEventDestroy(g->ev_fin);
	if (g->ev_copy) cudaEventDestroy(g->ev_copy);
	cudaFree(g->d_result);
	for (auto &p : g->ev_pending) {
		cudaEventDestroy(p.first);
		cudaEventDestroy(p.second);
	}
	for (auto &e : g->ev_pool)
		cudaEventDestroy(e);
	delete g;
}

/* re-send the tables and per-energy constants from pinned host memory (end-to-end timing: inputs travel every call) */
extern "C" int pcb_gpu_optic_reupload(pcb_gpu_optic *g, void *stream_v)
{
	cudaStream_t stream = (cudaStream_t)stream_v;
	CK(cudaSetDevice(g->device));
	CK(cudaMemcpyAsync(g->d_tables, g->h_tables, 5 * g->table_bytes, cudaMemcpyHostToDevice, stream));
	CK(cudaMemcpyAsync(g->d_ec, g->h_ec, g->ec_bytes, cudaMemcpyHostToDevice, stream));
	return 0;
}
extern "C" size_t pcb_gpu_optic_input_bytes(const pcb_gpu_optic *g) { return 5 * g->table_bytes + g->ec_bytes; }

/* grow-only pinned staging area owned by the optic (device -> host result copies run at PCIe speed from here) */
extern "C" void *pcb_gpu_optic_stage(pcb_gpu_optic *g, size_t bytes)
{
	if (bytes <= g->h_stage_bytes)
		return g->h_stage;
	cudaSetDevice(g->device);
	cudaFreeHost(g->h_stage);
	g->h_stage = nullptr;
	g->h_stage_bytes = 0;
	if (cudaHostAlloc(&g->h_stage, bytes, cudaHostAllocDefault) != cudaSuccess) {
		g_last_error = "cudaHostAlloc failed for the result staging area";
		g->h_stage = nullptr;
		return nullptr;
	}
	g->h_stage_bytes = bytes;
	return g->h_stage;
}

/* device memory for the results of one call, kept with the optic so that repeated calls do not cudaMalloc / cudaFree */
extern "C" void *pcb_gpu_optic_result(pcb_gpu_optic *g, size_t bytes)
{
	if (bytes <= g->d_result_bytes)
		return g->d_result;
	cudaSetDevice(g->device);
	cudaFree(g->d_result);
	g->d_result = nullptr;
	g->d_result_bytes = 0;
	if (cudaMalloc(&g->d_result, bytes) != cudaSuccess) {
		g_last_error = "cudaMalloc failed for the exit records";
		g->d_result = nullptr;
		return nullptr;
	}
	g->d_result_bytes = bytes;
	return g->d_result;
}

extern "C" double pcb_gpu_optic_hint(const pcb_gpu_optic *g, uint64_t key, int64_t *exits)
{
	if (g->hint_key != key || g->hint_exits <= 0)
		return 0.0;
	if (exits)
		*exits = g->hint_exits;
	return g->hint_frac;
}
extern "C" void pcb_gpu_optic_set_hint(pcb_gpu_optic *g, uint64_t key, double frac, int64_t exits)
{
	g->hint_key = key;
	g->hint_frac = frac;
	g->hint_exits = exits;
}

/* ---- instrumentation ---- */
extern "C" void pcb_gpu_optic_timing(pcb_gpu_optic *g, int on) { g->timing = on; }
extern "C" int64_t pcb_gpu_optic_kernels(const pcb_gpu_optic *g, int64_t *n_trace)
{
	if (n_trace)
		*n_trace = g->n_trace;
	return g->n_kernels;
}
/* total device time of the trace kernels launched since the last call (ms) and how many; waits for them */
extern "C" double pcb_gpu_optic_trace_ms(pcb_gpu_optic *g, int64_t *n_launches)
{
	cudaSetDevice(g->device);
	for (auto &p : g->ev_pending) {
		float ms = 0;
		cudaEventSynchronize(p.second);
		if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) {
			g->trace_ms += ms;
			g->trace_timed++;
		}
		g->ev_pool.push_back(p.first);
		g->ev_pool.push_back(p.second);
	}
	g->ev_pending.clear();
	double t = g->trace_ms;
	if (n_launches)
		*n_launches = g->trace_timed;
	g->trace_ms = 0;
	g->trace_timed = 0;
	return t;
}
static cudaEvent_t take_event(pcb_gpu_optic *g)
{
	cudaEvent_t e;
	if (!g->ev_pool.empty()) {
		e = g->ev_pool.back();
		g->ev_pool.pop_back();
		return e;
	}
	cudaEventCreate(&e);
	return e;
}

extern "C" int pcb_gpu_optic_device(const pcb_gpu_optic *g) { return g->device; }
extern "C" int pcb_gpu_optic_n_energies(const pcb_gpu_optic *g) { return g->nE; }

extern "C" void pcb_gpu_optic_tune(pcb_gpu_optic *g, int thr_cand, int thr_refl, int thr_new, int ctas_per_sm, int all_slow)
{
	if (thr_cand >= 0) g->thr_cand = thr_cand;
	if (thr_refl >= 0) g->thr_refl = thr_refl;
	if (thr_new >= 0) g->thr_new = thr_new;
	if (ctas_per_sm > 0 && ctas_per_sm != g->ctas_per_sm) {
		g->ctas_per_sm = ctas_per_sm;
		cudaSetDevice(g->device);
		cudaFree(g->d_wscratch);
		g->d_wscratch = nullptr;
	}
	if (all_slow >= 0) g->o.all_slow = all_slow;
}

static const int TALLY_BLOCKS = 512;

/* deferred weights (pcb_trace2.cuh, task ENERGY): source-driven runs of the task-queue tracer with the cheap Fresnel term and a
 * long energy grid.  Replay (polycap_photon_launch) always updates weights reflection by reflection: it returns the state of
 * absorbed photons too. */
static bool defer_on(const pcb_gpu_optic *g)
{
	return g->kernel_version == 2 && !g->exact_weights && g->defer_min > 0 && g->nE >= g->defer_min;
}

static int ensure_wave(pcb_gpu_optic *g, int64_t n)
{
	if (!g->d_wscratch) {
		size_t slots = (size_t)g->sm_count * g->ctas_per_sm * g->threads;
		if (slots < (size_t)g->sm_count * T2_NP)
			slots = (size_t)g->sm_count * T2_NP;
		CK(cudaMalloc(&g->d_wscratch, slots * (size_t)g->nE * 8));
	}
	if (!g->d_hist && defer_on(g))
		CK(cudaMalloc(&g->d_hist, (size_t)g->sm_count * T2_NP * T2_HK * 4 * 8));
	if (n <= g->wave_cap)
		return 0;
	cudaFree(g->d_status); cudaFree(g->d_fields); cudaFree(g->d_irefl); cudaFree(g->d_weights); cudaFree(g->d_rank);
	g->d_status = nullptr; g->d_fields = nullptr; g->d_irefl = nullptr; g->d_weights = nullptr; g->d_rank = nullptr;
	g->wave_cap = 0;
	CK(cudaMalloc(&g->d_status, (size_t)n));
	CK(cudaMalloc(&g->d_fields, (size_t)n * 17 * 8));
	CK(cudaMalloc(&g->d_irefl, (size_t)n * 4));
	CK(cudaMalloc(&g->d_weights, (size_t)n * g->nE * 8));
	CK(cudaMalloc(&g->d_rank, (size_t)n * 8));
	if (!g->d_blocksum) {
		CK(cudaMalloc(&g->d_blocksum, (size_t)(1 << 16) * 8));
		CK(cudaMalloc(&g->d_cutinfo, 4 * 8));
		CK(cudaMalloc(&g->d_partial, (size_t)TALLY_BLOCKS * g->nE * 8));
		CK(cudaMalloc(&g->d_cpartial, (size_t)TALLY_BLOCKS * 8 * 8));
	}
	g->wave_cap = n;
	return 0;
}

static int launch_trace(pcb_gpu_optic *g, int mode, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n, const PcbReplayIn &rin,
			cudaStream_t stream)
{
	PcbTraceParams P{};
	P.o = g->o;
	if (s) {
		P.s.d_source = s->d_source; P.s.src_x = s->src_x; P.s.src_y = s->src_y; P.s.src_sigx = s->src_sigx; P.s.src_sigy = s->src_sigy;
		P.s.src_shiftx = s->src_shiftx; P.s.src_shifty = s->src_shifty; P.s.hor_pol = s->hor_pol;
	}
	P.s.seed = seed;
	P.rin = rin;
	P.out.status = g->d_status; P.out.fields = g->d_fields; P.out.irefl = g->d_irefl; P.out.weights = g->d_weights;
	P.wc = g->d_wc;
	P.wscratch = g->d_wscratch;
	P.hist = g->d_hist;
	P.hk = g->hk > 0 ? (g->hk < T2_HK ? g->hk : T2_HK) : (g->nE >= 128 ? T2_HK : 32);
	P.id0 = (unsigned long long)id0;
	P.n_ids = (unsigned long long)n;
	P.mode = mode;
	P.tables_in_smem = g->tables_in_smem;
	P.thr_cand = g->thr_cand; P.thr_refl = g->thr_refl; P.thr_new = g->thr_new;
	P.record_all = 0;
	P.s.cap0 = g->cap0; /* uniform-illumination branch, source.c:79-89 */
	P.s.ext0 = g->ext0;

	CK(cudaMemsetAsync(g->d_wc, 0, sizeof(PcbWorkCounters), stream));
	int blocks = g->sm_count * g->ctas_per_sm;
	int64_t need = (n + g->threads - 1) / g->threads;
	if (need < blocks)
		blocks = (int)(need > 0 ? need : 1);
	size_t smem = g->tables_in_smem ? 5 * g->table_bytes : 0;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (g->timing) {
		e0 = take_event(g);
		e1 = take_event(g);
		cudaEventRecord(e0, stream);
	}
	g->n_kernels++;
	g->n_trace++;
	/* EXACT_W = 1: per-energy Fresnel term in the reference's operation order (bit-level weights); 0: the cheap evaluation */
#define PCB_LAUNCH(K, ...)                                                                                      \
	do {                                                                                                    \
		CK(cudaFuncSetAttribute(K<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
		K<__VA_ARGS__><<<blocks, nthreads, smem, stream>>>(P);                                          \
	} while (0)
	int nthreads = g->threads;
	if (g->kernel_version == 2) {
		nthreads = T2_THREADS;
		smem = t2_smem_bytes(g->o.tstride);
		blocks = g->sm_count;
		int64_t need2 = (n + T2_NP - 1) / T2_NP;
		if (need2 < blocks)
			blocks = (int)(need2 > 0 ? need2 : 1);
		if (mode == MODE_SOURCE) {
			if (g->exact_weights) PCB_LAUNCH(pcb_trace2_kernel, MODE_SOURCE, 1, 0);
			else if (defer_on(g) && g->d_hist) PCB_LAUNCH(pcb_trace2_kernel, MODE_SOURCE, 0, 1);
			else PCB_LAUNCH(pcb_trace2_kernel, MODE_SOURCE, 0, 0);
		} else {
			if (g->exact_weights) PCB_LAUNCH(pcb_trace2_kernel, MODE_REPLAY, 1, 0); else PCB_LAUNCH(pcb_trace2_kernel, MODE_REPLAY, 0, 0);
		}
	} else if (mode == MODE_SOURCE) {
		if (g->exact_weights) PCB_LAUNCH(pcb_trace_kernel, MODE_SOURCE, 1); else PCB_LAUNCH(pcb_trace_kernel, MODE_SOURCE, 0);
	} else {
		if (g->exact_weights) PCB_LAUNCH(pcb_trace_kernel, MODE_REPLAY, 1); else PCB_LAUNCH(pcb_trace_kernel, MODE_REPLAY, 0);
	}
#undef PCB_LAUNCH
	if (g->timing) {
		cudaEventRecord(e1, stream);
		g->ev_pending.emplace_back(e0, e1);
	}
	CK(cudaGetLastError());
	return 0;
}

/* ------------------------------------------------------------------ replay */

extern "C" int pcb_gpu_replay(pcb_gpu_optic *g, int64_t n, const double *coords, const double *dir, const double *ev, int *ret,
			      double *out_state, int64_t *irefl, double *weights, int64_t *counters)
{
	if (n <= 0)
		return 0;
	CK(cudaSetDevice(g->device));
	if (ensure_wave(g, n))
		return -1;
	if (n > g->replay_cap) {
		cudaFree(g->d_rin);
		g->d_rin = nullptr;
		g->replay_cap = 0;
		CK(cudaMalloc(&g->d_rin, (size_t)n * 9 * 8));
		g->replay_cap = n;
	}
	cudaStream_t stream = 0;
	CK(cudaMemcpyAsync(g->d_rin, coords, (size_t)n * 24, cudaMemcpyHostToDevice, stream));
	CK(cudaMemcpyAsync(g->d_rin + 3 * n, dir, (size_t)n * 24, cudaMemcpyHostToDevice, stream));
	CK(cudaMemcpyAsync(g->d_rin + 6 * n, ev, (size_t)n * 24, cudaMemcpyHostToDevice, stream));
	PcbReplayIn rin{g->d_rin, g->d_rin + 3 * n, g->d_rin + 6 * n};
	if (launch_trace(g, MODE_REPLAY, nullptr, 0, 0, n, rin, stream))
		return -1;
	std::vector<int8_t> st((size_t)n);
	std::vector<int> ir((size_t)n);
	std::vector<double> f((size_t)n * 11);
	CK(cudaMemcpyAsync(st.data(), g->d_status, (size_t)n, cudaMemcpyDeviceToHost, stream));
	CK(cudaMemcpyAsync(ir.data(), g->d_irefl, (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
	CK(cudaMemcpyAsync(f.data(), g->d_fields, (size_t)n * 11 * 8, cudaMemcpyDeviceToHost, stream));
	CK(cudaMemcpyAsync(weights, g->d_weights, (size_t)n * g->nE * 8, cudaMemcpyDeviceToHost, stream));
	PcbWorkCounters wc;
	CK(cudaMemcpyAsync(&wc, g->d_wc, sizeof wc, cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	for (int64_t i = 0; i < n; i++) {
		ret[i] = st[i];
		irefl[i] = ir[i];
		for (int k = 0; k < 11; k++)
			out_state[i * 11 + k] = f[(size_t)k * n + i];
	}
	if (counters) {
		memset(counters, 0, PCB_C_COUNT * 8);
		counters[PCB_C_IDS] = n;
		counters[PCB_C_REFL_ALL] = (int64_t)wc.reflections;
		counters[PCB_C_SEG_STEPS] = (int64_t)wc.seg_steps;
		counters[PCB_C_SOLVES] = (int64_t)wc.solves;
		counters[PCB_C_LAUNCHED] = (int64_t)wc.launched;
	}
	return 0;
}

/* ------------------------------------------------------------------ start states of ids (parity tooling) */

__global__ void sample_ids_kernel(PcbSource s, int mono, unsigned long long id0, long long n, double *out)
{
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	StartState st;
	source_sample(s, mono, id0 + (unsigned long long)i, st);
	double *o = out + i * 11;
	o[0] = st.coords.x; o[1] = st.coords.y; o[2] = st.coords.z;
	o[3] = st.dir.x; o[4] = st.dir.y; o[5] = st.dir.z;
	o[6] = st.ev.x; o[7] = st.ev.y; o[8] = st.ev.z;
	o[9] = st.src_x; o[10] = st.src_y;
}

extern "C" int pcb_gpu_sample_ids(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n, double *out_host)
{
	if (n <= 0)
		return 0;
	CK(cudaSetDevice(g->device));
	PcbSource ps{};
	ps.d_source = s->d_source; ps.src_x = s->src_x; ps.src_y = s->src_y; ps.src_sigx = s->src_sigx; ps.src_sigy = s->src_sigy;
	ps.src_shiftx = s->src_shiftx; ps.src_shifty = s->src_shifty; ps.hor_pol = s->hor_pol;
	ps.cap0 = g->cap0; ps.ext0 = g->ext0; ps.seed = seed;
	double *d = nullptr;
	CK(cudaMalloc(&d, (size_t)n * 11 * 8));
	sample_ids_kernel<<<(unsigned)((n + 127) / 128), 128>>>(ps, g->o.mono, (unsigned long long)id0, n, d);
	g->n_kernels++;
	cudaError_t e = cudaMemcpy(out_host, d, (size_t)n * 11 * 8, cudaMemcpyDeviceToHost);
	cudaFree(d);
	CK(e);
	return 0;
}

/* ------------------------------------------------------------------ wave post-processing */

static const int SCAN_T = 1024;

__global__ void wave_scan_blocks(const int8_t *status, int64_t n, int64_t *rank, int64_t *blocksum)
{
	__shared__ int warp_tot[32];
	int64_t i = (int64_t)blockIdx.x * SCAN_T + threadIdx.x;
	int flag = (i < n && status[i] == ST_EXIT) ? 1 : 0;
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	int v = flag;
	for (int off = 1; off < 32; off <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, v, off);
		if (lane >= off)
			v += t;
	}
	if (lane == 31)
		warp_tot[w] = v;
	__syncthreads();
	if (w == 0) {
		int t = warp_tot[lane];
		for (int off = 1; off < 32; off <<= 1) {
			int u = __shfl_up_sync(0xffffffffu, t, off);
			if (lane >= off)
				t += u;
		}
		warp_tot[lane] = t;
	}
	__syncthreads();
	int before = (w > 0 ? warp_tot[w - 1] : 0) + v - flag;
	if (i < n)
		rank[i] = before;
	if (threadIdx.x == SCAN_T - 1)
		blocksum[blockIdx.x] = before + flag;
}

/* single block: exclusive scan of block sums (<= 65536), total in cutinfo[2] */
__global__ void wave_scan_sums(int64_t *blocksum, int nb, int64_t *cutinfo)
{
	__shared__ int64_t carry;
	__shared__ int64_t wtot[32];
	if (threadIdx.x == 0)
		carry = 0;
	__syncthreads();
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int base = 0; base < nb; base += 1024) {
		int i = base + threadIdx.x;
		int64_t x = (i < nb) ? blocksum[i] : 0, v = x;
		for (int off = 1; off < 32; off <<= 1) {
			int64_t t = __shfl_up_sync(0xffffffffu, v, off);
			if (lane >= off)
				v += t;
		}
		if (lane == 31)
			wtot[w] = v;
		__syncthreads();
		if (w == 0) {
			int64_t t = wtot[lane];
			for (int off = 1; off < 32; off <<= 1) {
				int64_t u = __shfl_up_sync(0xffffffffu, t, off);
				if (lane >= off)
					t += u;
			}
			wtot[lane] = t;
		}
		__syncthreads();
		int64_t excl = carry + (w > 0 ? wtot[w - 1] : 0) + v - x;
		if (i < nb)
			blocksum[i] = excl;
		__syncthreads();
		if (threadIdx.x == 1023)
			carry = excl + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		cutinfo[2] = carry;
		cutinfo[0] = -1;
	}
}

/* adds block offsets; finds the id just after the n_wanted-th exit */
__global__ void wave_scan_finish(const int8_t *status, int64_t n, int64_t *rank, const int64_t *blocksum, int64_t n_wanted, int64_t *cutinfo)
{
	int64_t i = (int64_t)blockIdx.x * SCAN_T + threadIdx.x;
	if (i >= n)
		return;
	int64_t r = rank[i] + blocksum[blockIdx.x];
	rank[i] = r;
	if (n_wanted > 0 && status[i] == ST_EXIT && r == n_wanted - 1)
		cutinfo[0] = i + 1;
}

__global__ void wave_cut_resolve(int64_t n, int64_t n_wanted, int64_t *cutinfo)
{
	if (cutinfo[0] < 0) {
		cutinfo[0] = n;
		cutinfo[1] = cutinfo[2];
	} else {
		cutinfo[1] = n_wanted;
	}
}

/* fixed-order partial sums: block b owns ids [b*chunk, (b+1)*chunk) below the cut */
__global__ void wave_tally_partial(const int8_t *status, const int *irefl, const double *weights, int nE, int64_t n, const int64_t *cutinfo,
				   double *partial, int64_t *cpartial)
{
	extern __shared__ double sh[];
	const int64_t cut = cutinfo[0];
	const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
	int64_t lo = (int64_t)blockIdx.x * chunk, hi = lo + chunk;
	if (hi > cut)
		hi = cut;
	int EW = 1;
	while (EW < nE && EW < (int)blockDim.x)
		EW <<= 1;
	const int S = blockDim.x / EW;
	const int el = threadIdx.x % EW, sub = threadIdx.x / EW;
	for (int e0 = 0; e0 < nE; e0 += EW) {
		int e = e0 + el;
		double acc = 0;
		if (e < nE)
			for (int64_t i = lo + sub; i < hi; i += S)
				if (status[i] == ST_EXIT)
					acc += weights[i * nE + e];
		sh[threadIdx.x] = acc;
		__syncthreads();
		if (sub == 0 && e < nE) {
			double t = 0;
			for (int s = 0; s < S; s++)
				t += sh[s * EW + el];
			partial[(size_t)blockIdx.x * nE + e] = t;
		}
		__syncthreads();
	}
	/* counters: exit, not entered, absorbed, uncounted, refl of exit photons */
	long long c[5] = {0, 0, 0, 0, 0};
	for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
		int s = status[i];
		if (s == ST_EXIT) {
			c[0]++;
			c[4] += irefl[i];
		} else if (s == ST_NOT_ENTERED)
			c[1]++;
		else if (s == ST_ABSORBED)
			c[2]++;
		else
			c[3]++;
	}
	__shared__ long long cs[5];
	if (threadIdx.x < 5)
		cs[threadIdx.x] = 0;
	__syncthreads();
	for (int k = 0; k < 5; k++) {
		long long v = c[k];
		for (int off = 16; off > 0; off >>= 1)
			v += __shfl_down_sync(0xffffffffu, v, off);
		if ((threadIdx.x & 31) == 0)
			atomicAdd((unsigned long long *)&cs[k], (unsigned long long)v);
	}
	__syncthreads();
	if (threadIdx.x < 5)
		cpartial[(size_t)blockIdx.x * 8 + threadIdx.x] = cs[threadIdx.x];
}

__global__ void wave_tally_final(const double *partial, const int64_t *cpartial, int nb, int nE, const int64_t *cutinfo, const PcbWorkCounters *wc,
				 double *sum_weights, int64_t *counters)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < nE) {
		double acc = 0;
		for (int b = 0; b < nb; b++)
			acc += partial[(size_t)b * nE + t];
		sum_weights[t] += acc;
	}
	if (t < 5) {
		long long acc = 0;
		for (int b = 0; b < nb; b++)
			acc += cpartial[(size_t)b * 8 + t];
		const int map[5] = {PCB_C_EXIT, PCB_C_NOT_ENTERED, PCB_C_NOT_TRANSMITTED, PCB_C_UNCOUNTED, PCB_C_REFL_EXIT};
		counters[map[t]] += acc;
	}
	if (t == 5) {
		counters[PCB_C_IDS] += cutinfo[0];
		counters[PCB_C_REFL_ALL] += (int64_t)wc->reflections;
		counters[PCB_C_SEG_STEPS] += (int64_t)wc->seg_steps;
		counters[PCB_C_SOLVES] += (int64_t)wc->solves;
		counters[PCB_C_LAUNCHED] += (int64_t)wc->launched;
	}
	if (t >= 10 && t < 16)
		counters[t] += (int64_t)wc->dbg[t - 10]; /* scheduler statistics of the task-queue tracer: batches, lanes, empty polls, lost races, HIT batches, HIT lanes */
}

/* packed records: out_fields[f][rec_base + rank], out_weights[rec_base + rank][e] */
__global__ void wave_compact_fields(const int8_t *status, const int64_t *rank, const double *fields, int64_t n, const int64_t *cutinfo,
				    double *out_fields, int64_t rec_cap, int64_t rec_base)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cutinfo[0] || status[i] != ST_EXIT)
		return;
	int64_t r = rec_base + rank[i];
	if (r >= rec_cap)
		return;
#pragma unroll
	for (int f = 0; f < 17; f++)
		out_fields[(size_t)f * rec_cap + r] = fields[(size_t)f * n + i];
}

__global__ void wave_compact_weights(const int8_t *status, const int64_t *rank, const double *weights, int nE, int64_t n, const int64_t *cutinfo,
				     double *out_weights, int64_t rec_cap, int64_t rec_base)
{
	const int lane = threadIdx.x & 31;
	int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	const int64_t cut = cutinfo[0];
	for (int64_t i = warp; i < cut; i += nwarps) {
		if (status[i] != ST_EXIT)
			continue;
		int64_t r = rec_base + rank[i];
		if (r >= rec_cap)
			continue;
		for (int e = lane; e < nE; e += 32)
			out_weights[(size_t)r * nE + e] = weights[(size_t)i * nE + e];
	}
}

/* id-ordered post-processing of the dense per-id outputs in g->d_status/d_irefl/d_fields/d_weights (see file header), in two
 * steps so that a run over several devices can place the cut between them: wave_scan counts the wave's exit photons
 * (cutinfo[2]) and ranks them; wave_finish cuts after the n_wanted-th of them, tallies and compacts what lies below the cut. */
static int wave_scan(pcb_gpu_optic *g, int64_t n_ids, cudaStream_t stream)
{
	g->last_wave_n = n_ids;
	const int nb = (int)((n_ids + SCAN_T - 1) / SCAN_T);
	if (nb > (1 << 16)) {
		g_last_error = "wave too large (max 64M ids)";
		return -1;
	}
	wave_scan_blocks<<<nb, SCAN_T, 0, stream>>>(g->d_status, n_ids, g->d_rank, g->d_blocksum);
	wave_scan_sums<<<1, 1024, 0, stream>>>(g->d_blocksum, nb, g->d_cutinfo);
	g->n_kernels += 2;
	CK(cudaGetLastError());
	return 0;
}

static int wave_finish(pcb_gpu_optic *g, int64_t n_ids, int64_t n_wanted, double *d_sum_weights, int64_t *d_counters, double *d_rec_fields,
		       double *d_rec_weights, int64_t rec_cap, int64_t rec_base, cudaStream_t stream, int64_t *ids_used, int64_t *n_exit)
{
	const int nb = (int)((n_ids + SCAN_T - 1) / SCAN_T);
	wave_scan_finish<<<nb, SCAN_T, 0, stream>>>(g->d_status, n_ids, g->d_rank, g->d_blocksum, n_wanted, g->d_cutinfo);
	wave_cut_resolve<<<1, 1, 0, stream>>>(n_ids, n_wanted, g->d_cutinfo);
	int tb = TALLY_BLOCKS;
	if (tb > nb)
		tb = nb;
	wave_tally_partial<<<tb, 256, 256 * 8, stream>>>(g->d_status, g->d_irefl, g->d_weights, g->nE, n_ids, g->d_cutinfo, g->d_partial, g->d_cpartial);
	int ft = g->nE > 16 ? g->nE : 16;
	wave_tally_final<<<(ft + 255) / 256, 256, 0, stream>>>(g->d_partial, g->d_cpartial, tb, g->nE, g->d_cutinfo, g->d_wc, d_sum_weights, d_counters);
	g->n_kernels += 4;
	if (d_rec_fields && d_rec_weights) {
		g->n_kernels += 2;
		wave_compact_fields<<<(unsigned)((n_ids + 255) / 256), 256, 0, stream>>>(g->d_status, g->d_rank, g->d_fields, n_ids, g->d_cutinfo, d_rec_fields,
											  rec_cap, rec_base);
		wave_compact_weights<<<g->sm_count * 4, 256, 0, stream>>>(g->d_status, g->d_rank, g->d_weights, g->nE, n_ids, g->d_cutinfo, d_rec_weights, rec_cap,
									  rec_base);
	}
	CK(cudaGetLastError());
	if (ids_used || n_exit) {
		int64_t ci[3];
		CK(cudaMemcpyAsync(ci, g->d_cutinfo, sizeof ci, cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
		if (ids_used)
			*ids_used = ci[0];
		if (n_exit)
			*n_exit = ci[1];
	}
	return 0;
}

extern "C" int pcb_gpu_run_wave(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n_ids, int64_t n_wanted,
				double *d_sum_weights, int64_t *d_counters, double *d_rec_fields, double *d_rec_weights, int64_t rec_cap,
				int64_t rec_base, void *stream_v, int64_t *ids_used, int64_t *n_exit)
{
	cudaStream_t stream = (cudaStream_t)stream_v;
	if (n_ids <= 0)
		return 0;
	CK(cudaSetDevice(g->device));
	if (ensure_wave(g, n_ids))
		return -1;
	PcbReplayIn rin{};
	if (launch_trace(g, MODE_SOURCE, s, seed, id0, n_ids, rin, stream) || wave_scan(g, n_ids, stream))
		return -1;
	return wave_finish(g, n_ids, n_wanted, d_sum_weights, d_counters, d_rec_fields, d_rec_weights, rec_cap, rec_base, stream, ids_used, n_exit);
}

/* ---- the same wave in two asynchronous phases on the optic's own stream (pcb_api.cpp drives one optic per device) ---- */
static int ensure_streams(pcb_gpu_optic *g)
{
	if (g->s_main)
		return 0;
	/* blocking streams: they keep the implicit ordering with the legacy default stream, which the single-wave entry points
	 * (replay, run_ids, table re-upload) use; main and copy stream do not order against each other */
	CK(cudaStreamCreate(&g->s_main));
	CK(cudaStreamCreate(&g->s_copy));
	CK(cudaEventCreateWithFlags(&g->ev_fin, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&g->ev_copy, cudaEventDisableTiming));
	CK(cudaHostAlloc((void **)&g->h_pin, 8 * sizeof(int64_t), cudaHostAllocPortable));
	return 0;
}

extern "C" void *pcb_gpu_optic_stream(pcb_gpu_optic *g)
{
	if (cudaSetDevice(g->device) != cudaSuccess || ensure_streams(g))
		return nullptr;
	return (void *)g->s_main;
}

/* phase 1: trace ids [id0, id0+n_ids), rank the exit photons; their number travels to pinned memory (pcb_gpu_wave_count) */
extern "C" int pcb_gpu_wave_begin(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n_ids)
{
	if (n_ids <= 0)
		return 0;
	CK(cudaSetDevice(g->device));
	if (ensure_streams(g) || ensure_wave(g, n_ids))
		return -1;
	PcbReplayIn rin{};
	if (launch_trace(g, MODE_SOURCE, s, seed, id0, n_ids, rin, g->s_main) || wave_scan(g, n_ids, g->s_main))
		return -1;
	CK(cudaMemcpyAsync(g->h_pin + 2, g->d_cutinfo + 2, 8, cudaMemcpyDeviceToHost, g->s_main));
	return 0;
}

/* waits for phase 1: exit photons of the whole wave */
extern "C" int pcb_gpu_wave_count(pcb_gpu_optic *g, int64_t *exits)
{
	CK(cudaSetDevice(g->device));
	CK(cudaStreamSynchronize(g->s_main));
	*exits = g->h_pin[2];
	return 0;
}

/* phase 2 (asynchronous): cut after the n_wanted-th exit (0: none), tallies into d_sum_weights / d_counters, records appended at
 * rec_base of the device record buffer.  pcb_gpu_wave_result waits and returns ids below the cut / exits below the cut. */
extern "C" int pcb_gpu_wave_finish(pcb_gpu_optic *g, int64_t n_ids, int64_t n_wanted, double *d_sum_weights, int64_t *d_counters, double *d_rec_fields,
				   double *d_rec_weights, int64_t rec_cap, int64_t rec_base)
{
	CK(cudaSetDevice(g->device));
	if (wave_finish(g, n_ids, n_wanted, d_sum_weights, d_counters, d_rec_fields, d_rec_weights, rec_cap, rec_base, g->s_main, nullptr, nullptr))
		return -1;
	CK(cudaMemcpyAsync(g->h_pin, g->d_cutinfo, 16, cudaMemcpyDeviceToHost, g->s_main));
	CK(cudaEventRecord(g->ev_fin, g->s_main));
	return 0;
}

extern "C" int pcb_gpu_wave_result(pcb_gpu_optic *g, int64_t *ids_used, int64_t *n_exit)
{
	CK(cudaSetDevice(g->device));
	CK(cudaStreamSynchronize(g->s_main));
	*ids_used = g->h_pin[0];
	*n_exit = g->h_pin[1];
	return 0;
}

/* records [rec_base, rec_base + n) of the device record buffer -> host arrays laid out [17][host_cap] / [host_cap][nE], at record
 * host_base, on the optic's copy stream once phase 2 of the wave that produced them has finished (it overlaps the next trace) */
extern "C" int pcb_gpu_wave_download(pcb_gpu_optic *g, const double *d_rec_fields, const double *d_rec_weights, int64_t rec_cap, int64_t rec_base, int64_t n,
				     double *h_fields, double *h_weights, int64_t host_cap, int64_t host_base)
{
	if (n <= 0)
		return 0;
	CK(cudaSetDevice(g->device));
	CK(cudaStreamWaitEvent(g->s_copy, g->ev_fin, 0));
	CK(cudaMemcpy2DAsync(h_fields + host_base, (size_t)host_cap * 8, d_rec_fields + rec_base, (size_t)rec_cap * 8, (size_t)n * 8, 17, cudaMemcpyDeviceToHost,
			     g->s_copy));
	CK(cudaMemcpyAsync(h_weights + (size_t)host_base * g->nE, d_rec_weights + (size_t)rec_base * g->nE, (size_t)n * g->nE * 8, cudaMemcpyDeviceToHost,
			   g->s_copy));
	CK(cudaEventRecord(g->ev_copy, g->s_copy));
	return 0;
}

/* the next phase 2 may overwrite the record buffer from its start: make the main stream wait for the downloads queued so far */
extern "C" int pcb_gpu_wave_reuse_records(pcb_gpu_optic *g)
{
	CK(cudaSetDevice(g->device));
	CK(cudaStreamWaitEvent(g->s_main, g->ev_copy, 0));
	return 0;
}

extern "C" int pcb_gpu_wave_download_sync(pcb_gpu_optic *g)
{
	CK(cudaSetDevice(g->device));
	if (g->s_copy)
		CK(cudaStreamSynchronize(g->s_copy));
	return 0;
}

/* ---- tallies of several devices combined on one of them over peer memory (NVLink): out[i] = sum over devices, fixed order ---- */
struct PcbPeerPtrs {
	const double *sum[16];
	const int64_t *cnt[16];
	int n;
};
__global__ void tally_reduce_peers(PcbPeerPtrs pp, int nE, double *out_sum, int64_t *out_cnt)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < nE) {
		double acc = 0;
		for (int d = 0; d < pp.n; d++)
			acc += pp.sum[d][t];
		out_sum[t] = acc;
	}
	if (t < PCB_C_COUNT) {
		int64_t acc = 0;
		for (int d = 0; d < pp.n; d++)
			acc += pp.cnt[d][t];
		out_cnt[t] = acc;
	}
}

/* Sum of the per-device tallies (device pointers d_sum[k], d_cnt[k] on the devices of optics[k]) -> host.  With peer access the
 * first device reads the others' buffers directly (one kernel, one download); without it the buffers are staged through the host.
 * Either way the terms are added in device order, so the result does not depend on the path.  *used_peer reports which. */
extern "C" int pcb_gpu_reduce_tallies(pcb_gpu_optic *const *optics, double *const *d_sum, int64_t *const *d_cnt, int n, double *h_sum, int64_t *h_cnt,
				      int *used_peer)
{
	if (n < 1 || n > 16) {
		g_last_error = "pcb_gpu_reduce_tallies: 1..16 devices";
		return -1;
	}
	pcb_gpu_optic *g0 = optics[0];
	const int nE = g0->nE;
	bool peer = n > 1 && !getenv("POLYCAP_B200_NO_PEER");
	for (int k = 1; k < n && peer; k++) {
		int can = 0;
		if (optics[k]->device == g0->device)
			continue;
		if (cudaDeviceCanAccessPeer(&can, g0->device, optics[k]->device) != cudaSuccess || !can)
			peer = false;
	}
	if (used_peer)
		*used_peer = peer ? 1 : 0;
	for (int k = 0; k < n; k++) { /* every device's waves are complete */
		CK(cudaSetDevice(optics[k]->device));
		if (optics[k]->s_main)
			CK(cudaStreamSynchronize(optics[k]->s_main));
	}
	CK(cudaSetDevice(g0->device));
	if (n == 1) {
		CK(cudaMemcpy(h_sum, d_sum[0], (size_t)nE * 8, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(h_cnt, d_cnt[0], PCB_C_COUNT * 8, cudaMemcpyDeviceToHost));
		return 0;
	}
	if (peer) {
		for (int k = 1; k < n; k++)
			if (optics[k]->device != g0->device) {
				cudaError_t e = cudaDeviceEnablePeerAccess(optics[k]->device, 0);
				if (e == cudaErrorPeerAccessAlreadyEnabled)
					cudaGetLastError();
				else if (e != cudaSuccess) {
					cudaGetLastError();
					peer = false;
					if (used_peer)
						*used_peer = 0;
					break;
				}
			}
	}
	if (peer) {
		PcbPeerPtrs pp{};
		pp.n = n;
		for (int k = 0; k < n; k++) {
			pp.sum[k] = d_sum[k];
			pp.cnt[k] = d_cnt[k];
		}
		if (!g0->d_reduce)
			CK(cudaMalloc(&g0->d_reduce, (size_t)nE * 8 + PCB_C_COUNT * 8));
		double *out = g0->d_reduce;
		int64_t *out_c = (int64_t *)(out + nE);
		const int nt = nE > PCB_C_COUNT ? nE : PCB_C_COUNT;
		tally_reduce_peers<<<(nt + 255) / 256, 256, 0, g0->s_main>>>(pp, nE, out, out_c);
		g0->n_kernels++;
		CK(cudaMemcpyAsync(h_sum, out, (size_t)nE * 8, cudaMemcpyDeviceToHost, g0->s_main));
		CK(cudaMemcpyAsync(h_cnt, out_c, PCB_C_COUNT * 8, cudaMemcpyDeviceToHost, g0->s_main));
		CK(cudaStreamSynchronize(g0->s_main));
		return 0;
	}
	std::vector<double> ts((size_t)nE);
	int64_t tc[PCB_C_COUNT];
	for (int i = 0; i < nE; i++)
		h_sum[i] = 0;
	for (int i = 0; i < PCB_C_COUNT; i++)
		h_cnt[i] = 0;
	for (int k = 0; k < n; k++) {
		CK(cudaSetDevice(optics[k]->device));
		CK(cudaMemcpy(ts.data(), d_sum[k], (size_t)nE * 8, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(tc, d_cnt[k], sizeof tc, cudaMemcpyDeviceToHost));
		for (int i = 0; i < nE; i++)
			h_sum[i] += ts[i];
		for (int i = 0; i < PCB_C_COUNT; i++)
			h_cnt[i] += tc[i];
	}
	return 0;
}

/* ------------------------------------------------------------------ leak_calc = true (pcb_leak.cu) */

static void leak_view(const pcb_gpu_optic *g, pcb_leak_optic *v)
{
	v->nmax = g->nmax; v->nE = g->nE; v->mono = g->o.mono; v->monotonic = g->monotonic; v->sm_count = g->sm_count; v->device = g->device;
	v->n_shells = g->o.n_shells; v->sig_rough = g->o.sig_rough; v->cap0 = g->cap0; v->ext0 = g->ext0;
	v->d_z = g->o.z; v->d_cap = g->o.cap; v->d_ext = g->o.ext; v->d_hexz = g->o.hexz; v->d_ec = g->o.ec;
	const size_t tn = g->table_bytes / 8;
	v->d_se = g->d_leak_tables; v->d_sc = g->d_leak_tables + tn; v->d_sl = g->d_leak_tables + 2 * tn; v->d_isl = g->d_leak_tables + 3 * tn;
}

extern "C" int pcb_gpu_replay_leak(pcb_gpu_optic *g, int64_t n, const double *coords, const double *dir, const double *ev, int *ret, double *out_state,
				   int64_t *irefl, double *weights, int64_t *counters, int64_t *leak_count, double **leak_rec, int64_t *n_leak_rec)
{
	*leak_rec = nullptr;
	*n_leak_rec = 0;
	if (n <= 0)
		return 0;
	CK(cudaSetDevice(g->device));
	pcb_leak_optic v;
	leak_view(g, &v);
	std::vector<int8_t> st((size_t)n);
	std::vector<int> ir((size_t)n), ln((size_t)n);
	std::vector<double> f((size_t)n * 11);
	int64_t work[4] = {0, 0, 0, 0};
	if (pcb_leak_run(&v, nullptr, 0, 0, n, coords, dir, ev, st.data(), ir.data(), f.data(), weights, ln.data(), leak_rec, n_leak_rec, work)) {
		g_last_error = pcb_leak_last_error();
		return -1;
	}
	g->n_kernels++;
	const size_t stride = 11 + (size_t)g->nE;
	size_t r = 0;
	for (int64_t i = 0; i < n; i++) {
		ret[i] = st[i];
		irefl[i] = ir[i];
		for (int k = 0; k < 11; k++)
			out_state[i * 11 + k] = f[(size_t)k * n + i];
		if (leak_count) {
			leak_count[2 * i] = leak_count[2 * i + 1] = 0;
			for (int k = 0; k < ln[i]; k++, r++)
				leak_count[2 * i + ((*leak_rec)[r * stride] == 0. ? 0 : 1)]++;
		}
	}
	if (counters) {
		memset(counters, 0, PCB_C_COUNT * 8);
		counters[PCB_C_IDS] = n;
		counters[PCB_C_REFL_ALL] = work[0];
		counters[PCB_C_LAUNCHED] = n;
		counters[PCB_C_WALL_STEPS] = work[1];
		counters[PCB_C_WALL_EVALS] = work[2];
		counters[PCB_C_SOLVES] = work[3];
	}
	return 0;
}

extern "C" int pcb_gpu_run_wave_leak(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n_ids, int64_t n_wanted,
				     double *d_sum_weights, int64_t *d_counters, double *d_rec_fields, double *d_rec_weights, int64_t rec_cap,
				     int64_t rec_base, int64_t *ids_used, int64_t *n_exit, double **leak_rec, int64_t *n_leak_rec)
{
	*leak_rec = nullptr;
	*n_leak_rec = 0;
	if (n_ids <= 0)
		return 0;
	CK(cudaSetDevice(g->device));
	if (ensure_wave(g, n_ids))
		return -1;
	pcb_leak_optic v;
	leak_view(g, &v);
	const size_t nE = (size_t)g->nE;
	std::vector<int8_t> st((size_t)n_ids);
	std::vector<int> ir((size_t)n_ids), ln((size_t)n_ids);
	std::vector<double> f((size_t)n_ids * 17), w((size_t)n_ids * nE);
	double *recs = nullptr;
	int64_t nrec = 0, work[4] = {0, 0, 0, 0};
	if (int lrc = pcb_leak_run(&v, s, seed, id0, n_ids, nullptr, nullptr, nullptr, st.data(), ir.data(), f.data(), w.data(), ln.data(), &recs, &nrec, work)) {
		g_last_error = pcb_leak_last_error();
		return lrc; /* -2: record pool full, the caller retries with a smaller wave */
	}
	g->n_kernels++;
	/* same id-ordered post-processing as the plain path, over the leak kernel's dense outputs */
	cudaStream_t stream = 0;
	PcbWorkCounters wc{};
	wc.reflections = (unsigned long long)work[0];
	wc.launched = (unsigned long long)n_ids;
	wc.dbg[5] = (unsigned long long)work[1]; /* -> counters[PCB_C_WALL_STEPS] */
	wc.dbg[4] = (unsigned long long)work[2]; /* -> counters[PCB_C_WALL_EVALS] */
	bool ok = cudaMemcpy(g->d_status, st.data(), (size_t)n_ids, cudaMemcpyHostToDevice) == cudaSuccess &&
		  cudaMemcpy(g->d_irefl, ir.data(), (size_t)n_ids * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
		  cudaMemcpy(g->d_fields, f.data(), f.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
		  cudaMemcpy(g->d_weights, w.data(), w.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
		  cudaMemcpy(g->d_wc, &wc, sizeof wc, cudaMemcpyHostToDevice) == cudaSuccess;
	if (!ok) {
		free(recs);
		g_last_error = "leak wave: device upload failed";
		return -1;
	}
	int64_t used = 0, got = 0;
	if (wave_scan(g, n_ids, stream) ||
	    wave_finish(g, n_ids, n_wanted, d_sum_weights, d_counters, d_rec_fields, d_rec_weights, rec_cap, rec_base, stream, &used, &got)) {
		free(recs);
		return -1;
	}
	if (ids_used)
		*ids_used = used;
	if (n_exit)
		*n_exit = got;
	/* leak records of the ids below the cut */
	int64_t keep = 0;
	for (int64_t i = 0; i < used; i++)
		keep += ln[(size_t)i];
	*leak_rec = recs;
	*n_leak_rec = keep;
	return 0;
}

extern "C" int pcb_gpu_leak_unit_step(pcb_gpu_optic *g, int mode, const double in[16], int ix, double *weights, double outs[13], double **leak_rec,
				      int64_t *n_leak_rec)
{
	CK(cudaSetDevice(g->device));
	pcb_leak_optic v;
	leak_view(g, &v);
	g->n_kernels++;
	int rc = pcb_leak_unit_step(&v, mode, in, ix, weights, outs, leak_rec, n_leak_rec);
	if (rc)
		g_last_error = pcb_leak_last_error();
	return rc;
}

extern "C" int pcb_gpu_leak_trace_wall(pcb_gpu_optic *g, const double in[6], double *d_travel, int *r_cntr, int *q_cntr)
{
	CK(cudaSetDevice(g->device));
	pcb_leak_optic v;
	leak_view(g, &v);
	int rc = pcb_leak_unit_trace_wall(&v, in, d_travel, r_cntr, q_cntr);
	if (rc == -1 && *pcb_leak_last_error())
		g_last_error = pcb_leak_last_error();
	return rc;
}

extern "C" int pcb_gpu_last_wave_dense(pcb_gpu_optic *g, int64_t n_ids, int8_t *status, double *rec)
{
	CK(cudaSetDevice(g->device));
	if (n_ids > g->last_wave_n) {
		g_last_error = "no such wave";
		return -1;
	}
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(status, g->d_status, (size_t)n_ids, cudaMemcpyDeviceToHost));
	if (rec) {
		const int64_t n = g->last_wave_n;
		std::vector<double> f((size_t)n * 17), w((size_t)n_ids * g->nE);
		CK(cudaMemcpy(f.data(), g->d_fields, f.size() * 8, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(w.data(), g->d_weights, w.size() * 8, cudaMemcpyDeviceToHost));
		const size_t stride = 17 + g->nE;
		for (int64_t i = 0; i < n_ids; i++) {
			for (int k = 0; k < 17; k++)
				rec[i * stride + k] = f[(size_t)k * n + i];
			memcpy(rec + i * stride + 17, w.data() + (size_t)i * g->nE, (size_t)g->nE * 8);
		}
	}
	return 0;
}

/* ------------------------------------------------------------------ small device helpers */

extern "C" void *pcb_gpu_malloc(int device, size_t bytes)
{
	void *p = nullptr;
	if (cudaSetDevice(pick_device(device)) != cudaSuccess || cudaMalloc(&p, bytes) != cudaSuccess) {
		g_last_error = "cudaMalloc failed";
		return nullptr;
	}
	return p;
}
extern "C" void pcb_gpu_free(int device, void *p)
{
	cudaSetDevice(pick_device(device));
	cudaFree(p);
}
extern "C" void *pcb_gpu_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
		g_last_error = "cudaHostAlloc failed for the result arrays";
		return nullptr;
	}
	return p;
}
extern "C" void pcb_gpu_host_free(void *p) { cudaFreeHost(p); }
extern "C" int pcb_gpu_memset(int device, void *p, int v, size_t bytes, void *stream)
{
	CK(cudaSetDevice(pick_device(device)));
	CK(cudaMemsetAsync(p, v, bytes, (cudaStream_t)stream));
	return 0;
}
extern "C" int pcb_gpu_copy_to_host(int device, void *dst, const void *src, size_t bytes, void *stream)
{
	CK(cudaSetDevice(pick_device(device)));
	CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	CK(cudaStreamSynchronize((cudaStream_t)stream));
	return 0;
}
extern "C" int pcb_gpu_copy_dev(int device, void *dst, const void *src, size_t bytes, void *stream)
{
	CK(cudaSetDevice(pick_device(device)));
	CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
	return 0;
}
extern "C" int pcb_gpu_sync(int device, void *stream)
{
	CK(cudaSetDevice(pick_device(device)));
	CK(cudaStreamSynchronize((cudaStream_t)stream));
	return 0;
}

/* ------------------------------------------------------------------ division self-test */

/* pcb_div(a, pcb_divisor(b)) against the compiler's a / b on pseudo-random operand pairs whose exponents cover
 * `span` binades either side of 1 (span >= 1100 reaches zeros, subnormals, infinities and NaNs); counts bit mismatches */
__global__ void div_selftest_kernel(unsigned long long n, unsigned long long seed, int span, unsigned long long *bad)
{
	unsigned long long local = 0;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
		Philox rng;
		rng.init(seed, i);
		double v[2];
		for (int k = 0; k < 2; k++) {
			rng.block((uint32_t)k);
			const unsigned long long m = ((unsigned long long)rng.cache[0] << 32 | rng.cache[1]) & 0x800fffffffffffffULL;
			const unsigned sel = rng.cache[2];
			long long e = 1023 + (long long)(sel % (2u * span + 1u)) - span;
			e = e < 0 ? 0 : (e > 2047 ? 2047 : e);
			unsigned long long bits = m | ((unsigned long long)e << 52);
			if ((sel >> 28) == 0)
				bits &= 0xfff0000000000000ULL; /* powers of two, zeros, infinities */
			if ((sel >> 28) == 1)
				bits |= 0x000fffffffffffffULL; /* all-ones significands */
			if (k == 0 && (sel >> 28) == 2)
				bits &= 0x8000000000000000ULL; /* zero numerators of either sign */
			v[k] = __longlong_as_double((long long)bits);
		}
		const PcbDivisor d = pcb_divisor(v[1]);
		const double q = pcb_div(v[0], d), r = v[0] / v[1];
		const bool same = (__double_as_longlong(q) == __double_as_longlong(r)) || (q != q && r != r);
		if (!same)
			local++;
	}
	if (local)
		atomicAdd(bad, local);
}

extern "C" long long pcb_gpu_div_selftest(int device, long long n, unsigned long long seed, int span)
{
	device = pick_device(device);
	if (cudaSetDevice(device) != cudaSuccess)
		return -1;
	unsigned long long *d = nullptr, h = 0;
	if (cudaMalloc(&d, 8) != cudaSuccess)
		return -1;
	cudaMemset(d, 0, 8);
	div_selftest_kernel<<<1184, 256>>>((unsigned long long)n, seed, span, d);
	const cudaError_t rc = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
	cudaFree(d);
	return rc == cudaSuccess ? (long long)h : -1;
}

/* ------------------------------------------------------------------ FP64 peak */

__global__ void dfma_peak_kernel(double *out, int iters)
{
	double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double m = 1.0000001, c = 1e-9;
	for (int i = 0; i < iters; i++) {
		a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
		a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
	}
	out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" double pcb_gpu_dfma_peak(int device, int iters)
{
	device = pick_device(device);
	if (cudaSetDevice(device) != cudaSuccess)
		return -1;
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	int blocks = prop.multiProcessorCount * 8, threads = 256;
	double *d = nullptr;
	if (cudaMalloc(&d, (size_t)blocks * threads * 8) != cudaSuccess)
		return -1;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	dfma_peak_kernel<<<blocks, threads>>>(d, iters / 4 + 1);
	double best = 0;
	for (int rep = 0; rep < 3; rep++) {
		cudaEventRecord(e0);
		dfma_peak_kernel<<<blocks, threads>>>(d, iters);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
		if (tf > best)
			best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	return best;
}
