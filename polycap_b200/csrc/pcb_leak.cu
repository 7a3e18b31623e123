/* pcb_leak.cu -- leak/recap wall transport (placeholder until the leak kernel lands): every entry fails loudly. */
#include <cstdint>
#include <cstddef>
#include "pcb_gpu.h"
extern "C" int pcb_gpu_replay_leak(pcb_gpu_optic *, int64_t, const double *, const double *, const double *, int *, double *, int64_t *, double *, int64_t *,
				   int64_t *, double **, int64_t *) { return -1; }
extern "C" int pcb_gpu_run_wave_leak(pcb_gpu_optic *, const pcb_source_desc *, uint64_t, int64_t, int64_t, int64_t, double *, int64_t *, double *, double *,
				     int64_t, int64_t, int64_t *, int64_t *, double **, int64_t *) { return -1; }
extern "C" int pcb_unit_reflect_leak(const double *, const double *, const double *, const double *, int, int64_t, const double *, int, double, double *, double *,
				     int64_t *, double **, int64_t *) { return -100; }
extern "C" int pcb_unit_trace_leak(const double *, int *, long long *, const double *, const double *, const double *, const double *, const double *, int, int64_t,
				   const double *, int, double, double *, double *, int64_t *, double **, int64_t *) { return -100; }
extern "C" int pcb_unit_trace_wall(const double *, const double *, const double *, const double *, int, int64_t, double *, int *, int *) { return -100; }
