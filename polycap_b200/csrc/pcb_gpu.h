/*
 * pcb_gpu.h -- host-side handle on the device-resident optic and the kernel launchers.
 * Internal to libpolycap_b200 (the C ABI is include/polycap.h + include/polycap_b200.h).
 */
#ifndef PCB_GPU_H
#define PCB_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcb_gpu_optic pcb_gpu_optic;

typedef struct {
	int nmax;
	const double *z, *cap, *ext; /* nmax+1 each */
	double sig_rough;
	int64_t n_cap;
	double density;
	size_t n_energies;
	const double *energies, *amu, *scatf;
} pcb_optic_desc;

typedef struct {
	double d_source, src_x, src_y, src_sigx, src_sigy, src_shiftx, src_shifty, hor_pol;
} pcb_source_desc;

/* accumulated over ids, all int64 so one NCCL all-reduce covers them */
enum {
	PCB_C_IDS = 0,          /* photon ids consumed */
	PCB_C_EXIT,             /* reached the exit window */
	PCB_C_NOT_ENTERED,      /* hit the glass at the entrance (launch code 2) */
	PCB_C_NOT_TRANSMITTED,  /* absorbed (launch code 0) */
	PCB_C_UNCOUNTED,        /* -2, -1 and "returned 1 but missed the exit window" */
	PCB_C_REFL_EXIT,        /* sum of i_refl over exit photons (what the reference prints) */
	PCB_C_REFL_ALL,         /* capil_reflect events of all photons */
	PCB_C_SEG_STEPS,        /* segments visited */
	PCB_C_SOLVES,           /* full quadratic solves */
	PCB_C_LAUNCHED,
	/* 10..14: scheduler statistics of the task-queue tracer (batches, lanes, empty polls, lost races, HIT batches) */
	PCB_C_WALL_EVALS = 14,  /* leak runs (the slot is a scheduler statistic otherwise): wall-march steps actually evaluated, the rest was jumped over */
	PCB_C_WALL_STEPS = 15,  /* leak runs: steps of the wall march (capil.c:1018-1068) */
	PCB_C_COUNT = 16
};

const char *pcb_gpu_last_error(void);
int pcb_gpu_device_count(void);
/* the device a request for `device` lands on (-1: env POLYCAP_B200_DEVICE, else the calling thread's current device) */
int pcb_gpu_resolve_device(int device);

/* Uploads profile tables and per-energy optical constants to `device` (-1: current default). */
pcb_gpu_optic *pcb_gpu_optic_create(int device, const pcb_optic_desc *d);
void pcb_gpu_optic_destroy(pcb_gpu_optic *g);
int pcb_gpu_optic_device(const pcb_gpu_optic *g);
int pcb_gpu_optic_n_energies(const pcb_gpu_optic *g);

/* Tuning knobs (phase-vote thresholds in lanes, CTAs per SM, exact-check mode); <0 keeps the default. */
void pcb_gpu_optic_tune(pcb_gpu_optic *g, int thr_cand, int thr_refl, int thr_new, int ctas_per_sm, int all_slow);

/* pinned-memory plumbing: re-send the optic's inputs (H2D), bytes of those inputs, grow-only pinned staging area for results */
int pcb_gpu_optic_reupload(pcb_gpu_optic *g, void *stream);
size_t pcb_gpu_optic_input_bytes(const pcb_gpu_optic *g);
void *pcb_gpu_optic_stage(pcb_gpu_optic *g, size_t bytes);
void *pcb_gpu_optic_result(pcb_gpu_optic *g, size_t bytes);
double pcb_gpu_optic_hint(const pcb_gpu_optic *g, uint64_t key, int64_t *exits);
void pcb_gpu_optic_set_hint(pcb_gpu_optic *g, uint64_t key, double frac, int64_t exits);

/* instrumentation (bench.py): CUDA-event timing of the trace kernels, kernel-launch counts */
void pcb_gpu_optic_timing(pcb_gpu_optic *g, int on);
int64_t pcb_gpu_optic_kernels(const pcb_gpu_optic *g, int64_t *n_trace);
double pcb_gpu_optic_trace_ms(pcb_gpu_optic *g, int64_t *n_launches);

/* Replay: n start states in host memory -> per-photon results in host memory.
 * out_state: [n][11] = exit coords(3), exit dir(3), exit elecv(3), d_travel, capil_trace calls */
int pcb_gpu_replay(pcb_gpu_optic *g, int64_t n, const double *coords, const double *dir, const double *ev, int *ret, double *out_state,
		   int64_t *irefl, double *weights, int64_t *counters /* [PCB_C_COUNT] or NULL */);

/* One source-driven wave over ids [id0, id0+n_ids) on `stream` (a cudaStream_t, NULL = default).
 * Device-resident accumulation: d_sum_weights[nE] += exit weights, d_counters[PCB_C_COUNT] += ...
 * n_wanted > 0 cuts the wave after its n_wanted-th exit photon (ids beyond the cut are not consumed).
 * If rec_fields/rec_weights (device, capacity rec_cap records) are given, exit records are appended in
 * id order at rec_base.  Returns the number of ids consumed via *ids_used and exits via *n_exit
 * (this synchronises the stream when either pointer or n_wanted is given). */
int pcb_gpu_run_wave(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n_ids, int64_t n_wanted,
		     double *d_sum_weights, int64_t *d_counters, double *d_rec_fields, double *d_rec_weights, int64_t rec_cap, int64_t rec_base,
		     void *stream, int64_t *ids_used, int64_t *n_exit);

/* The same wave in two asynchronous phases on the optic's own stream, so that a run over several devices can place the cut
 * between them and record downloads overlap the next trace (pcb_api.cpp: polycap_source_get_transmission_efficiencies).
 *   begin     trace ids [id0, id0+n_ids) and rank the exit photons          wave_count     wait; exits of the whole wave
 *   finish    cut after the n_wanted-th exit (0: none), tally, compact      wave_result    wait; ids / exits below the cut
 *   download  records -> host arrays [17][host_cap], [host_cap][nE] on the copy stream, after that wave's finish
 *   reuse_records  the next finish may overwrite the record buffer from its start (waits for the downloads queued so far) */
void *pcb_gpu_optic_stream(pcb_gpu_optic *g);
int pcb_gpu_wave_begin(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n_ids);
int pcb_gpu_wave_count(pcb_gpu_optic *g, int64_t *exits);
int pcb_gpu_wave_finish(pcb_gpu_optic *g, int64_t n_ids, int64_t n_wanted, double *d_sum_weights, int64_t *d_counters, double *d_rec_fields,
			double *d_rec_weights, int64_t rec_cap, int64_t rec_base);
int pcb_gpu_wave_result(pcb_gpu_optic *g, int64_t *ids_used, int64_t *n_exit);
int pcb_gpu_wave_download(pcb_gpu_optic *g, const double *d_rec_fields, const double *d_rec_weights, int64_t rec_cap, int64_t rec_base, int64_t n,
			  double *h_fields, double *h_weights, int64_t host_cap, int64_t host_base);
int pcb_gpu_wave_reuse_records(pcb_gpu_optic *g);
int pcb_gpu_wave_download_sync(pcb_gpu_optic *g);
/* sum over n devices of their tallies -> host (peer reads over NVLink when available, else staged; fixed device order) */
int pcb_gpu_reduce_tallies(pcb_gpu_optic *const *optics, double *const *d_sum, int64_t *const *d_cnt, int n, double *h_sum, int64_t *h_cnt, int *used_peer);

/* start state the tracer draws for ids [id0, id0+n): out[n][11] = coords(3) dir(3) elecv(3) src_x src_y (host memory) */
int pcb_gpu_sample_ids(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n, double *out_host);

/* leak_calc = true (pcb_leak.cu): replay / source wave with leak records, and the single-item wall march.
 * leak_rec: malloc'd [n_leak_rec][11+nE] = kind (0 ext, 1 int), coords(3), direction(3), elecv(3), n_refl, weights(nE), photon order. */
int pcb_gpu_replay_leak(pcb_gpu_optic *g, int64_t n, const double *coords, const double *dir, const double *ev, int *ret, double *out_state, int64_t *irefl,
			double *weights, int64_t *counters, int64_t *leak_count /* [n][2] */, double **leak_rec, int64_t *n_leak_rec);
int pcb_gpu_run_wave_leak(pcb_gpu_optic *g, const pcb_source_desc *s, uint64_t seed, int64_t id0, int64_t n_ids, int64_t n_wanted, double *d_sum_weights,
			  int64_t *d_counters, double *d_rec_fields, double *d_rec_weights, int64_t rec_cap, int64_t rec_base, int64_t *ids_used,
			  int64_t *n_exit, double **leak_rec, int64_t *n_leak_rec);
/* one polycap_capil_reflect (mode 1) / polycap_capil_trace (mode 2) with leak_calc = true; arguments as pcb_leak_unit_step (pcb_leak.h) */
int pcb_gpu_leak_unit_step(pcb_gpu_optic *g, int mode, const double in[16], int ix, double *weights, double outs[13], double **leak_rec,
			   int64_t *n_leak_rec);
int pcb_gpu_leak_trace_wall(pcb_gpu_optic *g, const double in[6], double *d_travel, int *r_cntr, int *q_cntr);

/* dense per-id view of the last wave (debug / parity): status[n] and the 17+nE record of every id */
int pcb_gpu_last_wave_dense(pcb_gpu_optic *g, int64_t n_ids, int8_t *status, double *rec /* [n][17+nE] or NULL */);

/* device memory helpers for callers that do not link the CUDA runtime */
void *pcb_gpu_malloc(int device, size_t bytes);
void pcb_gpu_free(int device, void *p);
void *pcb_gpu_host_alloc(size_t bytes); /* pinned, portable */
void pcb_gpu_host_free(void *p);
int pcb_gpu_memset(int device, void *p, int v, size_t bytes, void *stream);
int pcb_gpu_copy_to_host(int device, void *dst, const void *src_dev, size_t bytes, void *stream);
int pcb_gpu_sync(int device, void *stream);
int pcb_gpu_copy_dev(int device, void *dst_dev, const void *src_dev, size_t bytes, void *stream); /* asynchronous, same device */

/* FP64 FMA throughput of the device in TFLOP/s (2 flops per DFMA), timed with CUDA events */
double pcb_gpu_dfma_peak(int device, int iters);
long long pcb_gpu_div_selftest(int device, long long n, unsigned long long seed, int span);

#ifdef __cplusplus
}
#endif
#endif
