/*
 * include/polycap_b200.h -- additive, GPU-specific entry points of libpolycap_b200.so.
 *
 * Everything in include/polycap.h keeps the reference's semantics.  The functions here expose
 * what the reference API cannot express: batched replay of many photons in one kernel launch,
 * runs addressed by photon id (Philox4x32-10 stream per id, so any GPU can take any id range),
 * tally-only runs beyond what `int n_photons` x n_energies can materialise, and device-resident
 * tallies that a caller (e.g. torch.distributed / NCCL) can all-reduce without a host round trip.
 *
 * C ABI only: plain pointers and sizes, no CUDA or torch types in any signature
 * (streams and device pointers travel as void* / double* / int64_t*).
 */
#ifndef POLYCAP_B200_H
#define POLYCAP_B200_H

#include "polycap.h"

#ifdef __cplusplus
extern "C" {
#endif

/* layout of the int64 counter block (host or device), one all-reduce covers it */
enum {
	POLYCAP_B200_C_IDS = 0,         /* photon ids consumed (every call of the sampler) */
	POLYCAP_B200_C_EXIT,            /* photons that reached the exit window (reference: iexit) */
	POLYCAP_B200_C_NOT_ENTERED,     /* hit the glass at the entrance (reference: not_entered) */
	POLYCAP_B200_C_NOT_TRANSMITTED, /* absorbed inside (reference: not_transmitted) */
	POLYCAP_B200_C_UNCOUNTED,       /* -2, -1, or returned 1 outside the exit window: in no reference counter */
	POLYCAP_B200_C_REFL_EXIT,       /* sum of i_refl over exit photons (what the reference prints) */
	POLYCAP_B200_C_REFL_ALL,        /* reflection events of all photons */
	POLYCAP_B200_C_SEG_STEPS,       /* capillary segments visited */
	POLYCAP_B200_C_SOLVES,          /* full ray/cone solves */
	POLYCAP_B200_C_LAUNCHED,
	/* 10..14: scheduler statistics of the tracer */
	POLYCAP_B200_C_WALL_STEPS = 15, /* leak_calc = true: steps of the wall march (src/polycap-capil.c:1018-1068) */
	POLYCAP_B200_N_COUNTERS = 16
};

/* number of usable CUDA devices (0 when there is no driver/device) and the last device-side error text */
POLYCAP_EXTERN int polycap_b200_device_count(void);
POLYCAP_EXTERN const char *polycap_b200_last_error(void);
POLYCAP_EXTERN const char *polycap_b200_optconst_provider(void);

/* amu[i] (1/cm) and scatf[i] exactly as the library derives them for `description`
 * (replaces the per-photon polycap_photon_scatf of src/polycap-photon.c:22-94). */
POLYCAP_EXTERN bool polycap_b200_optical_constants(polycap_description *description, size_t n_energies, const double *energies, double *amu,
						   double *scatf, polycap_error **error);

/* Batched polycap_photon_launch (src/polycap-photon.c:390-955): n photons, one launch.
 * Inputs  start_coords/start_direction/start_electric_vector: [n][3].
 * amu/scatf may be NULL (then the provider above is used) or [n_energies] arrays to pin them.
 * Outputs ret[n] (1/0/2/-2/-1), exit_coords/exit_direction/exit_electric_vector [n][3] (the state
 * the reference leaves in the photon struct: last wall hit, not the exit plane), i_refl[n],
 * d_travel[n], weights [n][n_energies].  Any output pointer may be NULL. */
POLYCAP_EXTERN bool polycap_b200_photon_launch_batch(polycap_description *description, size_t n_energies, const double *energies,
						     const double *amu, const double *scatf, int64_t n, const double *start_coords,
						     const double *start_direction, const double *start_electric_vector, bool leak_calc, int *ret,
						     double *exit_coords, double *exit_direction, double *exit_electric_vector, int64_t *i_refl,
						     double *d_travel, double *weights, int64_t *counters, polycap_error **error);

/* The same with leak_calc = true (leak / recap wall transport, src/polycap-capil.c:896-1194, 658-887).  leak_count[n][2] receives
 * the number of external and internal leak records of each photon; *leak_records is malloc'd (free with polycap_free) and holds
 * *n_leak_records records of 11 + n_energies doubles in photon order: kind (0 external, 1 internal), coords(3), direction(3),
 * electric vector(3), n_refl, weights(n_energies) -- the fields of polycap_leak (include/polycap-photon.h:52-59). */
POLYCAP_EXTERN bool polycap_b200_photon_launch_batch_leaks(polycap_description *description, size_t n_energies, const double *energies,
							   const double *amu, const double *scatf, int64_t n, const double *start_coords,
							   const double *start_direction, const double *start_electric_vector, int *ret, double *exit_coords,
							   double *exit_direction, double *exit_electric_vector, int64_t *i_refl, double *d_travel,
							   double *weights, int64_t *leak_count, double **leak_records, int64_t *n_leak_records,
							   polycap_error **error);

/* Source-driven run over photon ids [id0, id0+n_ids) with seed `seed`, results on the host:
 * sum_weights[nE] (exit photons only), counters[POLYCAP_B200_N_COUNTERS], optional status[n_ids]
 * (1 exit, 0 absorbed, 2 not entered, -2/-1, 3 = reached the end outside the exit window) and optional
 * rec[n_ids][17+nE]: for exit photons the 17 values of struct _polycap_images (src start x,y; start x,y;
 * start dir x,y; start elecv x,y; exit x,y,z; exit dir x,y; exit elecv x,y; n_refl; d_travel) + weights. */
POLYCAP_EXTERN bool polycap_b200_source_run_ids(polycap_source *source, uint64_t seed, int64_t id0, int64_t n_ids, bool leak_calc,
						double *sum_weights, int64_t *counters, int8_t *status, double *rec, polycap_error **error);

/* counters[POLYCAP_B200_N_COUNTERS] of the calling process's last polycap_source_get_transmission_efficiencies call (work counters for
 * benchmarks: reflections of all photons, segments visited, wall-march steps ...); false when there was none */
POLYCAP_EXTERN bool polycap_b200_last_run_counters(int64_t *counters);

/* The start state (source.c:23-144 over the id's Philox stream) the device draws for ids [id0, id0+n_ids):
 * out[n_ids][11] = start coords(3), start direction(3), electric vector(3), source point x, y. */
POLYCAP_EXTERN bool polycap_b200_source_sample_ids(polycap_source *source, uint64_t seed, int64_t id0, int64_t n_ids, double *out, polycap_error **error);

/* ---- sessions: device-resident tallies for benchmarks and multi-GPU reductions -------------- */
typedef struct _polycap_b200_session polycap_b200_session;

/* binds `source` to CUDA device `device` (-1: current / env POLYCAP_B200_DEVICE) and uploads its tables */
POLYCAP_EXTERN polycap_b200_session *polycap_b200_session_new(polycap_source *source, int device, uint64_t seed, polycap_error **error);
POLYCAP_EXTERN void polycap_b200_session_free(polycap_b200_session *session);
/* Enqueues, on `cuda_stream` (a cudaStream_t cast to void*, NULL = default stream), the trace of ids
 * [id0, id0+n_ids) in waves of at most wave_size ids and ACCUMULATES into the caller's DEVICE buffers
 * d_sum_weights[nE] (double) and d_counters[POLYCAP_B200_N_COUNTERS] (int64).  Asynchronous: returns
 * after the launches are queued.  Tally-only: no per-photon records are kept. */
POLYCAP_EXTERN bool polycap_b200_session_run(polycap_b200_session *session, int64_t id0, int64_t n_ids, int64_t wave_size, void *cuda_stream,
					     double *d_sum_weights, int64_t *d_counters, polycap_error **error);
/* phase-vote thresholds (lanes) and resident CTAs per SM of the tracer; negative keeps the current value */
POLYCAP_EXTERN void polycap_b200_session_tune(polycap_b200_session *session, int thr_cand, int thr_refl, int thr_new, int ctas_per_sm);

/* Instrumentation for benchmarks.  _timing(on): bracket every trace-kernel launch with CUDA events on its stream.
 * _trace_ms: waits for the timed launches since the last call, returns their summed device time [ms] and count.
 * _kernels: kernels this session has launched so far (all of them / trace kernels only). */
POLYCAP_EXTERN void polycap_b200_session_timing(polycap_b200_session *session, int on);
POLYCAP_EXTERN double polycap_b200_session_trace_ms(polycap_b200_session *session, int64_t *n_launches);
POLYCAP_EXTERN int64_t polycap_b200_session_kernels(polycap_b200_session *session, int64_t *n_trace_kernels);

/* ---- streams: the exit records of a run delivered chunk by chunk ------------------------------------------------------------
 * polycap_source_get_transmission_efficiencies materialises n_photons x n_energies weights in one result object and counts in
 * `int` (/root/reference/include/polycap-source.h:114-121, src/polycap-source.c:675): a 1e9-photon run with 291 energies does not
 * fit that shape.  A stream walks the same id-ordered photon sequence (seed, ids 0, 1, 2 ...) and hands the records over in
 * caller-sized chunks while the tallies keep accumulating on the device(s); POLYCAP_B200_DEVICES shards it like the plain call. */
typedef struct _polycap_b200_stream polycap_b200_stream;
POLYCAP_EXTERN polycap_b200_stream *polycap_b200_stream_new(polycap_source *source, uint64_t seed, polycap_error **error);
POLYCAP_EXTERN void polycap_b200_stream_free(polycap_b200_stream *stream);
/* Continues the run until max_records further photons reached the exit window (the id-ordered sequence is cut right after the last
 * of them) or max_ids further ids have been traced (0: no limit).  Their records go to fields[17][max_records] (the 17 values of
 * polycap_b200_source_run_ids, field-major) and weights[max_records][n_energies]; both NULL = tally only.  Returns the number of
 * records delivered (< max_records only when max_ids ran out), -1 on error. */
POLYCAP_EXTERN int64_t polycap_b200_stream_next(polycap_b200_stream *stream, int64_t max_records, int64_t max_ids, double *fields, double *weights,
						polycap_error **error);
/* running totals of the stream: photon ids consumed, sum of exit weights per energy, counters[POLYCAP_B200_N_COUNTERS] */
POLYCAP_EXTERN bool polycap_b200_stream_totals(polycap_b200_stream *stream, int64_t *ids_consumed, double *sum_weights, int64_t *counters,
					       polycap_error **error);

/* FP64 FMA throughput of `device` in TFLOP/s (2 flops per DFMA) -- the roofline denominator */
POLYCAP_EXTERN double polycap_b200_dfma_peak_tflops(int device, int iters);

/* Device self-test of the shared-divisor division the kernels use in place of repeated `a / b` (pcb_math.h): compares it
 * with the compiler's IEEE division on `n` pseudo-random operand pairs spanning `span` binades either side of 1
 * (>= 1100 includes zeros, subnormals, infinities, NaNs).  Returns the number of bit mismatches (0 expected), -1 on error. */
POLYCAP_EXTERN int64_t polycap_b200_selftest_division(int device, int64_t n, uint64_t seed, int span);

#ifdef __cplusplus
}
#endif
#endif
