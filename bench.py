#!/usr/bin/env python
"""bench.py -- photons/s and reflections/s of the photon-transport hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload conical|xos1|ellip|dubfoc|parab|monocap]
                    [--sweep] [--leaks] [--no-subs]

Workload (N=1 default, BASELINE.json configs[1] / SURVEY 8d C2): conical polycapillary
(POLYCAP_PROFILE_CONICAL, radii of example/cone.*), 200000 capillaries, SiO2, source at 100 cm,
10-energy grid 1..30 keV; one STEP = one pass of the hot path over a batch of 1e7 photon ids per GPU
(source sampling with Philox4x32-10 per id -> hexagonal lookup -> segment scan -> reflection loop with the
per-energy Fresnel weight update -> device-resident tallies).  "photons" are counted the way the reference
counts "Simulated photons" (iexit + not_entered + not_transmitted, /root/reference/src/polycap-source.c:1061).

  value      device-timed (CUDA events, max over ranks), everything resident in HBM, whole job over N GPUs
  e2e        the same metric through the reference-facing C-ABI call polycap_source_get_transmission_efficiencies
             with HOST buffers: table upload (H2D) and exit-record download (D2H) inside the timed region
  roofline   FP64-pipe bound: algorithmic flops (SURVEY 8d model x device work counters) / trace-kernel time
             (CUDA events on its stream) against an in-run DFMA peak measurement; an HBM line is given as well
  cpu_baseline / --impl reference: the reference's own OpenMP path (oracle/_ref, its sources compiled in place)
             on the box's host cores on a bounded sample of the same workload
"""
import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# xraylib is not in this image: both arms use the library's built-in optical-constants model, by explicit request (it is opt-in)
os.environ.setdefault("POLYCAP_B200_OPTCONST", "builtin")
os.environ.setdefault("POLYCAP_B200_QUIET", "1")

SIO2 = {8: 53.0, 14: 47.0}
WORKLOADS = {
    # name: (profile args, n_cap, sig_rough, source args, energies, ids per step per GPU)
    "conical": dict(profile=(0, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5), n_cap=200000, sig_rough=0.0,
                    source=(100.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.0), energies=np.linspace(1, 30, 10), ids=10_000_000,
                    desc="conical polycap (POLYCAP_PROFILE_CONICAL, cone.* radii), 1e7 photon ids/step/GPU, 10-energy grid 1-30 keV"),
    "ellip": dict(profile=(2, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5), n_cap=200000, sig_rough=0.0,
                  source=(2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.0), energies=np.linspace(1, 40, 400), ids=1_000_000,
                  desc="ellipsoidal polycap (ellip_l9 geometry), 1e6 photon ids/step/GPU, 400-energy grid 1-40 keV, leaks off"),
    # BASELINE configs[3]: the measured XOS profile (the z/cap/ext tables of example/xos1.{prf,axs,ext} as stored in the golden
    # fixture tests/golden/replay_xos1.npz), xos1.inp source and energy grid
    "xos1": dict(profile_arrays="xos1", n_cap=200000, sig_rough=0.0,
                 source=(2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.0), energies=1.0 + 0.1 * np.arange(291), ids=2_000_000,
                 desc="XOS measured profile (xos1.prf/.ext tables), 2e6 photon ids/step/GPU, 291-energy grid 1-30 keV (xos1.inp)"),
    # BASELINE configs[4]: example/dub_foc.inp -- ellipsoidal double-focusing optic, 227701 capillaries, rough walls
    # (5 A), 0.2 x 0.01 cm source at 500 cm with 0.24 mrad divergence, 90 % horizontal polarisation, 246 energies 1-50 keV
    "dubfoc": dict(profile=(2, 6., 0.2883, 0.07, 0.00035, 0.000085, 500., 0.25), n_cap=227701, sig_rough=5.0,
                   source=(500.0, 0.2, 0.01, 0.00024, 0.00024, 0.0, 0.0, 0.9), energies=1.0 + 0.2 * np.arange(246), ids=2_000_000,
                   desc="double-focusing polycap (example/dub_foc.inp: ellipsoidal, rough walls, divergent polarised source), 2e6 photon ids/step/GPU, 246-energy grid 1-50 keV"),
    # ... and the same radii with POLYCAP_PROFILE_PARABOLOIDAL (the config's name; dub_foc.inp itself says type 2)
    "parab": dict(profile=(1, 6., 0.2883, 0.07, 0.00035, 0.000085, 500., 0.25), n_cap=227701, sig_rough=5.0,
                  source=(500.0, 0.2, 0.01, 0.00024, 0.00024, 0.0, 0.0, 0.9), energies=1.0 + 0.2 * np.arange(246), ids=2_000_000,
                  desc="paraboloidal polycap with the dub_foc.inp radii and source, 2e6 photon ids/step/GPU, 246-energy grid 1-50 keV"),
    # BASELINE configs[0]: example/monocap.inp -- one conical capillary, uniform illumination, 80 energies; n_cap = 1 is
    # only reachable through the .inp parser (polycap_description_new refuses it), so the source is built from a generated .inp
    "monocap": dict(inp=True, profile=(0, 15., 0.012, 0.012, 0.01, 0.0005, 1000., 0.5), n_cap=1, sig_rough=0.0,
                    source=(5.0, 0.0001, 0.0001, -1.0, 0.0, 0.0, 0.0, 0.0), energies=1.0 + 1.0 * np.arange(80), e_grid=(1.0, 80.0, 1.0), ids=2_000_000,
                    desc="mono-capillary (example/monocap.inp via the .inp parser), 2e6 photon ids/step/GPU, 80-energy grid 1-80 keV"),
}
SUB_WORKLOADS = ("xos1", "ellip", "dubfoc", "monocap")  # short device-timed records added to the default line (driver-visible)


def inp_text(w, n_photons=1000):
    """the workload as a polycap .inp file (format of example/*.inp, src/polycap-source.c:261-446)"""
    s, p = w["source"], w["profile"]
    e0, e1, de = w["e_grid"]
    return ("%.17g\n%.17g\n%.17g %.17g\n%.17g %.17g\n%.17g %.17g\n%.17g\n2\n8 53.0\n14 47.0\n2.23\n%.17g %.17g %.17g\n%d\n%d\n"
            "%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n%d.\nbench.out\n"
            % (w["sig_rough"], s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], e0, e1, de, n_photons, p[0], *p[1:], w["n_cap"]))


def write_inp(w):
    import tempfile
    d = tempfile.mkdtemp(prefix="pcb_bench_")
    path = os.path.join(d, "w.inp")
    with open(path, "w") as f:
        f.write(inp_text(w))
    return path

# algorithmic FP64 flops of the minimal result-equivalent algorithm
F_LAUNCH, F_STEP, F_REFL, F_FRESNEL, F_EXIT = 155, 24, 445, 70, 100


def algorithmic_flops(cnt, nE):
    return (F_LAUNCH * cnt["launched"] + F_STEP * cnt["seg_steps"] + F_REFL * cnt["refl_all"] + F_FRESNEL * nE * cnt["refl_all"]
            + F_EXIT * cnt["exit"])


@contextlib.contextmanager
def quiet_stdout():
    """the C libraries printf their run summaries (source.c:1061-1063); keep stdout to the one JSON line"""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ------------------------------------------------------------------ reference arm / CPU baseline (oracle/_ref)
def reference_setup(w):
    from oracle import harness as H
    if not H.ref_available():
        raise RuntimeError("oracle/_ref/libpolycap_ref.so is missing (built by oracle/Makefile where /root/reference exists)")
    ref = H.Ref()
    L = ref.L
    if w.get("inp"):
        src = L.polycap_source_new_from_file(write_inp(w).encode(), None)
        if not src:
            raise RuntimeError("the reference could not parse the generated .inp")
        return ref, src
    if "profile_arrays" in w:
        g = np.load(os.path.join(ROOT, "tests", "golden", "replay_%s.npz" % w["profile_arrays"]))
        z, cap, ext = (np.ascontiguousarray(g[k], dtype=np.float64) for k in ("z", "cap", "ext"))
        prof = L.polycap_profile_new_from_arrays(int(g["nmax"]), ext.ctypes.data_as(H.c_double_p), cap.ctypes.data_as(H.c_double_p),
                                                 z.ctypes.data_as(H.c_double_p), None)
    else:
        prof = L.polycap_profile_new(*w["profile"], None)
    desc = ref.description(prof, w["sig_rough"], w["n_cap"], list(SIO2.keys()), list(SIO2.values()), 2.23)
    src = ref.source(desc, *w["source"], w["energies"])
    return ref, src


def reference_run(ref, src, n_exit, threads=-1):
    """one call of the reference's own polycap_source_get_transmission_efficiencies; returns (seconds, started, exits, reflections)"""
    with quiet_stdout():
        t0 = time.perf_counter()
        eff = ref.L.polycap_source_get_transmission_efficiencies(src, threads, int(n_exit), False, None, None)
        dt = time.perf_counter() - t0
    im = eff.contents.images.contents
    started, exits = im.i_start, im.i_exit
    refl = int(np.ctypeslib.as_array(im.pc_exit_nrefl, (exits,)).sum())
    nE = eff.contents.n_energies
    # Sum of exit weights per energy, recovered from the efficiencies the reference returns (source.c:1075: eff = sum_w / started)
    sum_w = np.ctypeslib.as_array(eff.contents.efficiencies, (nE,)).copy() * started
    ref.L.polycap_transmission_efficiencies_free(eff)
    return dt, started, exits, refl, sum_w


def reference_sample(w, seconds, steps=1, warmup=0):
    """bounded sample: n_exit sized from a probe so that one call takes about `seconds`"""
    ref, src = reference_setup(w)
    cores = os.cpu_count() or 1
    dt, started, exits, _, _ = reference_run(ref, src, 40)  # first call: thread start-up dominates, only good for an order of magnitude
    n1 = max(40, int(40 * 1.0 / max(dt, 1e-3)))
    dt, started, exits, _, _ = reference_run(ref, src, n1)  # ~1 s probe
    n = max(40, int(n1 * seconds / max(dt, 1e-3)))
    for _ in range(warmup):
        reference_run(ref, src, max(40, n // 4))
    tot_t = tot_started = tot_refl = tot_exit = 0
    tot_w = 0.0
    for _ in range(steps):
        dt, started, exits, refl, sum_w = reference_run(ref, src, n)
        tot_t += dt; tot_started += started; tot_refl += refl; tot_exit += exits
        tot_w = tot_w + sum_w
    ref.L.polycap_source_free(src)
    return dict(value=tot_started / tot_t, unit="photons/s", cores=cores, kind="reference",
                sample="%d call(s) of the reference's polycap_source_get_transmission_efficiencies (oracle/_ref: its own sources, gcc -O2 -fopenmp, "
                       "xraylib replaced by the built-in optical-constants table => per-photon scatf is cheaper than with real xraylib), "
                       "n_photons=%d exits/call = %d simulated photons in %.1f s on %d host threads" % (steps, n, tot_started, tot_t, cores),
                exit_photons_per_s=tot_exit / tot_t, reflections_of_exit_photons_per_s=tot_refl / tot_t, ms_per_step=1e3 * tot_t / steps,
                started=int(tot_started), exits=int(tot_exit), efficiencies=(tot_w / max(tot_started, 1)).tolist())


def stat_parity(eff_gpu, n_gpu, eff_ref, n_ref, exit_gpu=None, exit_ref=None):
    """north_star check 2: transmission efficiency per energy, GPU vs the reference's own CPU run of the same workload, in units of
    the binomial standard error at the two photon counts (independent random streams: Philox on the GPU, MT19937 in the reference).
    A photon's weight lies in [0, 1], so p(1-p)/N bounds the variance of the mean weight p."""
    eg, er = np.asarray(eff_gpu, dtype=np.float64), np.asarray(eff_ref, dtype=np.float64)
    pooled = (eg * n_gpu + er * n_ref) / max(n_gpu + n_ref, 1)
    sigma = np.sqrt(np.maximum(pooled * (1.0 - pooled), 1e-12) * (1.0 / max(n_gpu, 1) + 1.0 / max(n_ref, 1)))
    z = np.abs(eg - er) / sigma
    out = {"check": "efficiency per energy: |gpu - reference| / binomial sigma at the two photon counts; pass = all < 3",
           "n_sigma_max": float(z.max()), "n_sigma_rms": float(np.sqrt(np.mean(z * z))), "energies_over_3_sigma": int((z >= 3).sum()),
           "n_energies": int(z.size), "pass": bool((z < 3).all()), "photons_gpu": int(n_gpu), "photons_reference": int(n_ref),
           "max_abs_diff": float(np.abs(eg - er).max())}
    if exit_gpu is not None and exit_ref is not None:
        pg, pr = exit_gpu / max(n_gpu, 1), exit_ref / max(n_ref, 1)
        pp = (exit_gpu + exit_ref) / max(n_gpu + n_ref, 1)
        out["exit_fraction"] = {"gpu": pg, "reference": pr,
                                "n_sigma": float(abs(pg - pr) / np.sqrt(max(pp * (1 - pp), 1e-12) * (1.0 / max(n_gpu, 1) + 1.0 / max(n_ref, 1))))}
    return out


# ------------------------------------------------------------------ the product arm
class Ctx:
    """process-wide state of a bench run: rank layout, torch handles"""
    pass


def build_source(pc, w):
    if w.get("inp"):
        return pc.Source.new_from_file(write_inp(w))
    if "profile_arrays" in w:
        g = np.load(os.path.join(ROOT, "tests", "golden", "replay_%s.npz" % w["profile_arrays"]))
        prof = pc.Profile.new_from_arrays(g["ext"], g["cap"], g["z"])
    else:
        prof = pc.Profile(*w["profile"])
    desc = pc.Description(prof, w["sig_rough"], w["n_cap"], SIO2, 2.23)
    return pc.Source(desc, *w["source"], w["energies"])


def barrier(ctx):
    ctx.torch.cuda.synchronize()
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.torch.cuda.synchronize()


def measure_device(ctx, pc, src, nE, ids_per_step, steps, warmup, clocks=False, id_base=0):
    """`steps` timed passes of the hot path over fresh photon ids, inputs resident in HBM: CUDA events on the launching stream around
    every step, the all-reduce of the tallies (the one collective of the path) timed as part of the job, MAX over ranks"""
    torch, dist = ctx.torch, ctx.dist
    from polycap_b200.distributed import shard_range, allreduce_tallies
    ses = pc.Session(src, device=ctx.local_rank, seed=20000)
    ses.timing(True)
    t_sum = torch.zeros(nE, dtype=torch.float64, device=ctx.dev)
    t_cnt = torch.zeros(pc.N_COUNTERS, dtype=torch.int64, device=ctx.dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step(k):
        """one pass over this rank's shard of step k's ids (weak scaling: ids_per_step per GPU)"""
        lo, hi = shard_range(ids_per_step * ctx.world, ctx.rank, ctx.world, id0=id_base + k * ids_per_step * ctx.world)
        ses.run(lo, hi - lo, t_sum.data_ptr(), t_cnt.data_ptr(), wave_size=ids_per_step, stream=stream)

    for k in range(warmup):
        step(k)
    allreduce_tallies(t_sum, t_cnt)  # warm the collective too (first use sets up the communicator's channels)
    barrier(ctx)
    ses.trace_ms()  # drop warm-up timings
    k0 = ses.kernels()
    t_sum.zero_(); t_cnt.zero_()
    sampler = ClockSampler(ctx.local_rank) if (clocks and ctx.rank == 0) else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps + 1)]
    barrier(ctx)
    for k in range(steps):
        ctx.flush.fill_(k & 0xff)  # L2 flush, outside the event pair
        ev[k][0].record()
        step(warmup + k)
        ev[k][1].record()
    ev[steps][0].record()
    allreduce_tallies(t_sum, t_cnt)  # NCCL over NVLink when world > 1 (t includes the tally reduction)
    ev[steps][1].record()
    barrier(ctx)
    clk = sampler.stop() if sampler else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=ctx.dev)
    trace_ms, trace_launches = ses.trace_ms()
    trace = torch.tensor([trace_ms], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(trace, op=dist.ReduceOp.MAX)
    cnt = dict(zip(pc.COUNTER_NAMES, t_cnt.cpu().tolist()))
    k1 = ses.kernels()
    return dict(cnt=cnt, sw=t_sum.cpu().numpy(), total_s=float(total_ms.item()) * 1e-3, trace_s=float(trace.item()) * 1e-3, trace_launches=int(trace_launches),
                launches=int(k1[0] - k0[0]) + (2 if ctx.world > 1 else 0), clocks=clk, allreduce_ms=step_ms[steps])


def leak_counts(te):
    """i_extleak, i_intleak of the result's struct _polycap_images (src/polycap-private.h:156-181) -- read in place: the getters would
    build one object per leak record (millions at 400 energies)"""
    import ctypes as C
    P = C.c_void_p

    class _Im(C.Structure):
        _fields_ = [("i_start", C.c_int64), ("i_exit", C.c_int64), ("arrays", P * 18), ("i_extleak", C.c_int64), ("ext", P * 7), ("i_intleak", C.c_int64)]

    class _Te(C.Structure):
        _fields_ = [("n_energies", C.c_size_t), ("energies", P), ("efficiencies", P), ("images", C.POINTER(_Im))]
    im = C.cast(te._trans_eff, C.POINTER(_Te)).contents.images.contents
    return int(im.i_extleak), int(im.i_intleak)


def result_counts(te):
    import ctypes as C

    class _Im(C.Structure):
        _fields_ = [("i_start", C.c_int64), ("i_exit", C.c_int64)]

    class _Te(C.Structure):
        _fields_ = [("n_energies", C.c_size_t), ("energies", C.c_void_p), ("efficiencies", C.c_void_p), ("images", C.POINTER(_Im))]
    im = C.cast(te._trans_eff, C.POINTER(_Te)).contents.images.contents
    return int(im.i_start), int(im.i_exit)


def measure_e2e(ctx, pc, src, nE, n_exit, reps, devices=None, all_ranks=True, cold=False):
    """the same metric through the reference-facing call polycap_source_get_transmission_efficiencies with HOST buffers: tables and
    per-energy constants re-sent from pinned host memory every call, exit records returned in host arrays, wall clock around the
    calls.  devices: value for POLYCAP_B200_DEVICES (the library shards the call over those devices from this one process)."""
    torch, dist = ctx.torch, ctx.dist
    env = {"POLYCAP_B200_REUPLOAD": "1", "POLYCAP_B200_SEED": "777"}
    if devices is None:
        env["POLYCAP_B200_DEVICE"] = str(ctx.local_rank)
    else:
        env["POLYCAP_B200_DEVICES"] = devices
    saved = {k: os.environ.get(k) for k in list(env) + ["POLYCAP_B200_DEVICE", "POLYCAP_B200_DEVICES"]}
    os.environ.pop("POLYCAP_B200_DEVICES", None)
    os.environ.update(env)
    cold_s = None
    try:
        with quiet_stdout():
            # the first call of a process pays for device buffers, pinning the result block and a probe wave (no exit-fraction estimate yet)
            tc = time.perf_counter()
            first = src.get_transmission_efficiencies(-1, n_exit)
            cold_s = time.perf_counter() - tc
            cold_started = result_counts(first)[0]
            # warm-up at full size with two results alive at a time, as in the timed loop (the previous result is released only after
            # the next call returned: two pinned result blocks)
            warm = [first, src.get_transmission_efficiencies(-1, n_exit)]
            del warm, first
            if all_ranks:
                barrier(ctx)
            t0 = time.perf_counter()
            per_call = []
            last = None
            for _ in range(reps):
                tc = time.perf_counter()
                last = src.get_transmission_efficiencies(-1, n_exit)
                per_call.append(time.perf_counter() - tc)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    i_start, i_exit = result_counts(last)  # same seed => every call simulates the same photons
    del last
    t_e2e = torch.tensor([dt], dtype=torch.float64, device=ctx.dev)
    n_e2e = torch.tensor([float(i_start * reps)], dtype=torch.float64, device=ctx.dev)
    if all_ranks and ctx.world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_e2e, op=dist.ReduceOp.SUM)
    n_dev = 1 if devices is None else (pc.device_count() if devices == "all" else len(devices.split(",")))
    out = {"value": float(n_e2e.item() / t_e2e.item()), "unit": "photons/s",
           "h2d_bytes_per_step": int((5 * 8 * 1002 + 6 * 8 * nE) * n_dev), "d2h_bytes_per_step": int(n_exit * 8 * (17 + nE) + (8 * nE + 8 * pc.N_COUNTERS) * n_dev),
           "call": "polycap_source_get_transmission_efficiencies(source, -1, n_photons=%d exits) x %d, tables + per-energy constants re-sent from pinned host memory "
                   "per call (POLYCAP_B200_REUPLOAD), exit records returned in host arrays (downloaded wave by wave while the next wave is traced)" % (n_exit, reps),
           "simulated_photons_per_call": int(i_start), "seconds": float(t_e2e.item()), "seconds_per_call": [round(t, 4) for t in per_call],
           "record_download_GBps": n_exit * 8 * (17 + nE) / (dt / reps) / 1e9,
           "cold_first_call": {"seconds": round(cold_s, 4), "photons_per_s": cold_started / cold_s,
                               "note": "first call of the process for this source: device buffers, pinned result block, probe wave"}}
    if devices is not None:
        out["devices"] = devices
        out["n_devices"] = n_dev
    return out


def roofline(ctx, pc, m, nE, workload, kernel_name):
    cnt = m["cnt"]
    flops = algorithmic_flops(cnt, nE)
    trace_s, total_s, launches = m["trace_s"], m["total_s"], max(m["trace_launches"], 1)
    peaks = measured_peaks()
    hbm_peak = (peaks or {}).get("hbm_gbs", 6650.0)
    rec_bytes = cnt["exit"] * 8 * (17 + nE) * 2 + cnt["ids"] * (1 + 4 + 8)  # dense exit records written + re-read by the tally; status/irefl/rank per id
    # per GPU: rank 0's trace time over rank 0's share of the work (weak scaling: equal shares)
    achieved_tf = flops / ctx.world / trace_s / 1e12 if trace_s > 0 else 0.0
    # DRAM bytes per launch of the dominant kernel: per-id figure of the committed ncu capture x ids per launch (None without one)
    traffic, traffic_src = None, None
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", fn))).get(workload)
            if tr:
                traffic = tr["dram_bytes_per_id"] * cnt["ids"] / ctx.world / launches
                traffic_src = "ncu dram bytes per photon id (%s) x ids per launch" % tr["capture"]
                break
        except (OSError, ValueError, KeyError):
            pass
    return {"bound": "fp64", "achieved": achieved_tf, "peak": ctx.dfma_peak, "unit": "TFLOP/s", "frac": achieved_tf / ctx.dfma_peak if ctx.dfma_peak > 0 else None,
            "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name, "kernel_ms_per_launch": 1e3 * trace_s / launches,
            "kernel_share_of_step": trace_s / total_s if total_s > 0 else None, "algorithmic_flops_per_launch": flops / ctx.world / launches,
            "model": "155*launched + 24*seg_steps + 445*refl + 70*nE*refl + 100*exit (SURVEY 8d)",
            "peak_source": "DFMA microbenchmark run on this GPU inside bench.py (polycap_b200_dfma_peak_tflops); MEASURED_PEAKS.json holds no FP64 figure",
            "hbm": {"achieved": rec_bytes / ctx.world / trace_s / 1e9 if trace_s > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}}


def kernel_name(nE):
    if os.environ.get("POLYCAP_B200_KERNEL") == "1":
        return "pcb_trace_kernel<MODE_SOURCE, cheap weights>"
    defer_min = int(os.environ.get("POLYCAP_B200_DEFER_MIN", "48"))
    deferred = defer_min > 0 and nE >= defer_min
    return "pcb_trace2_kernel<MODE_SOURCE, cheap weights, %s>" % ("deferred per-energy products" if deferred else "immediate weight update")


def sub_record(ctx, pc, name, steps=3, warmup=2, ids=None):
    """a short device-timed record of another BASELINE workload inside the same line (same protocol, fewer steps)"""
    w = WORKLOADS[name]
    nE = len(w["energies"])
    src = build_source(pc, w)
    ids = ids or w["ids"]
    m = measure_device(ctx, pc, src, nE, ids, steps, warmup)
    cnt = m["cnt"]
    photons = cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"]
    rf = roofline(ctx, pc, m, nE, name, kernel_name(nE))
    return {"workload": w["desc"], "n_energies": nE, "steps": steps, "warmup": warmup, "photon_ids_per_step_per_gpu": ids, "n_gpus": ctx.world,
            "value": photons / m["total_s"], "unit": "photons/s", "photon_ids_per_s": cnt["ids"] / m["total_s"], "reflections_per_s": cnt["refl_all"] / m["total_s"],
            "exit_photons_per_s": cnt["exit"] / m["total_s"], "ms_per_step": 1e3 * m["total_s"] / steps, "roofline_frac": rf["frac"],
            "achieved_tflops": rf["achieved"], "kernel": rf["kernel"], "kernel_share_of_step": rf["kernel_share_of_step"],
            "efficiency_first_mid_last": [float((m["sw"] / max(photons, 1))[i]) for i in (0, nE // 2, nE - 1)]}


def leak_record(ctx, pc, w, n_exit, cpu_seconds):
    """BASELINE configs[2]: the run with leak / recap wall transport on (polycap_source_get_transmission_efficiencies(leak_calc=true)),
    end to end through the C ABI, beside the reference's own leak mode on the host cores"""
    nE = len(w["energies"])
    src = build_source(pc, w)
    os.environ["POLYCAP_B200_SEED"] = "4242"
    with quiet_stdout():
        src.get_transmission_efficiencies(-1, max(50, n_exit // 10), True)  # buffers
        t0 = time.perf_counter()
        te = src.get_transmission_efficiencies(-1, n_exit, True)
        dt = time.perf_counter() - t0
    started, exits = result_counts(te)
    n_ext, n_int = leak_counts(te)
    cnt = pc.last_run_counters() or {}
    # SURVEY 8d flop model with the leak terms: 76 instead of 70 per (reflection x energy), 60 per step of the wall march
    flops = (F_LAUNCH * cnt.get("launched", 0) + F_REFL * cnt.get("refl_all", 0) + 76 * nE * cnt.get("refl_all", 0) + F_EXIT * exits + 60 * cnt.get("wall_steps", 0))
    out = {"workload": w["desc"] + " -- leak_calc = true", "n_energies": nE, "value": started / dt, "unit": "photons/s", "seconds": dt, "exits": exits,
           "simulated_photons": started, "external_leak_records": n_ext, "internal_leak_records": n_int,
           "reflection_events_incl_secondaries": cnt.get("refl_all"), "wall_march_steps": cnt.get("wall_steps"),
           "wall_march_steps_evaluated": cnt.get("q_hit_batches"), "started_photons_per_s": started / dt, "exit_photons_per_s": exits / dt,
           "roofline": {"bound": "fp64", "achieved": flops / dt / 1e12, "peak": ctx.dfma_peak, "unit": "TFLOP/s", "frac": flops / dt / 1e12 / ctx.dfma_peak if ctx.dfma_peak else None,
                        "model": "155*launched + 445*refl + 76*nE*refl + 100*exit + 60*wall_steps (SURVEY 8d, leak terms; wall_steps = the steps of the reference's march, of which only wall_march_steps_evaluated are walked, the rest is jumped over exactly); segment steps of the leak tracer are not counted",
                        "kernel": "pcb_leak_kernel (one warp per primary photon: energies and segment scan dealt over the lanes, wall march jumped; explicit frame stack)"}}
    try:
        ref, rsrc = reference_setup(w)
        with quiet_stdout():
            t0 = time.perf_counter()
            n_ref = 0
            r_started = 0
            while time.perf_counter() - t0 < cpu_seconds:
                eff = ref.L.polycap_source_get_transmission_efficiencies(rsrc, -1, 8, True, None, None)
                r_started += eff.contents.images.contents.i_start
                n_ref += 8
                ref.L.polycap_transmission_efficiencies_free(eff)
            rdt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": r_started / rdt, "unit": "photons/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": "reference leak mode, %d exits = %d simulated photons in %.1f s" % (n_ref, r_started, rdt)}
    except Exception as exc:
        out["cpu_baseline"] = {"value": None, "sample": "unavailable: %s" % exc}
    return out


def sweep_c5(ctx, pc, cpu_seconds):
    """dub_foc.inp geometry and the same radii as a paraboloidal profile; sweeps of the source offset and of the source
    divergence; per point device-timed photons/s and the 3-sigma check of the efficiencies against the reference's CPU run"""
    points = [("src_shiftx", v) for v in (0.0, 0.05, 0.1, 0.2)] + [("sigma", v) for v in (1.2e-4, 2.4e-4, 4.8e-4, 9.6e-4)]
    out = []
    for name in ("dubfoc", "parab"):
        for what, v in points:
            w = dict(WORKLOADS[name])
            s = list(w["source"])
            if what == "src_shiftx":
                s[5] = v
            else:
                s[3] = s[4] = v
            w["source"] = tuple(s)
            nE = len(w["energies"])
            src = build_source(pc, w)
            m = measure_device(ctx, pc, src, nE, 1_000_000, 2, 1)
            cnt = m["cnt"]
            photons = cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"]
            rec = {"profile": name, what: v, "value": photons / m["total_s"], "unit": "photons/s", "reflections_per_s": cnt["refl_all"] / m["total_s"],
                   "exit_fraction": cnt["exit"] / max(cnt["ids"], 1), "roofline_frac": roofline(ctx, pc, m, nE, name, kernel_name(nE))["frac"]}
            if ctx.rank == 0 and cpu_seconds > 0:
                try:
                    b = reference_sample(w, cpu_seconds)
                    rec["cpu_baseline_photons_per_s"] = b["value"]
                    rec["stat_parity"] = stat_parity(m["sw"] / max(photons, 1), photons, b["efficiencies"], b["started"], cnt["exit"], b["exits"])
                except Exception as exc:
                    rec["stat_parity"] = {"unavailable": str(exc)}
            out.append(rec)
            print("sweep point", json.dumps(rec)[:300], file=sys.stderr)
    return out


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="conical", choices=sorted(WORKLOADS))
    ap.add_argument("--ids", type=int, default=0, help="photon ids per step per GPU (default: the workload's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="size of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-subs", action="store_true", help="skip the short records of the other BASELINE workloads")
    ap.add_argument("--sweep", action="store_true", help="SURVEY 8d C5: source-offset and divergence sweeps of the dub_foc / paraboloidal optics")
    ap.add_argument("--leaks", type=int, default=0, metavar="N_EXIT", help="add a leak_calc = true record of the workload with this many exit photons")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = WORKLOADS[args.workload]
    nE = len(w["energies"])
    ids_per_step = args.ids or w["ids"]
    config = {"workload": w["desc"] if not args.ids else w["desc"] + " [ids/step/GPU overridden: %d]" % ids_per_step,
              "n_energies": nE, "photon_ids_per_step_per_gpu": ids_per_step, "sharding": "contiguous photon-id ranges, 1 rank per GPU",
              "l2": "flushed between timed steps (256 MiB write); every step traces fresh ids"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is the OpenMP path on ALL host cores
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        steps = max(1, args.steps)
        per_step = max(3.0, min(20.0, 120.0 / (steps + 0.25 * args.warmup + 0.5)))
        b = reference_sample(w, per_step, steps=steps, warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "photons/s", "value": b["value"], "unit": "photons/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": b["value"], "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "exit_photons_per_s": b["exit_photons_per_s"], "reflections_of_exit_photons_per_s": b["reflections_of_exit_photons_per_s"]}
        print(json.dumps(line))
        return 0

    # Libraries below (NCCL's version banner, the C library's run summaries) write to fd 1; the contract is ONE JSON line on
    # stdout, so fd 1 is pointed at stderr for the duration and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import polycap_b200 as pc

    if not torch.cuda.is_available() or pc.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.rank, ctx.local_rank, ctx.world = torch, dist, rank, local_rank, world
    ctx.dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=ctx.dev)
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)
    ctx.dfma_peak = pc.dfma_peak_tflops(local_rank, 400000)

    src = build_source(pc, w)
    m = measure_device(ctx, pc, src, nE, ids_per_step, args.steps, args.warmup, clocks=True)
    cnt, sw, total_s = m["cnt"], m["sw"], m["total_s"]
    photons = cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"]

    # ---- e2e through the reference-facing call, host buffers, copies inside the timed region
    e2e = None
    exit_frac = max(cnt["exit"] / max(cnt["ids"], 1), 1e-6)
    n_exit = max(1000, int(exit_frac * ids_per_step))
    n_exit = min(n_exit, 2_000_000_000 // (8 * (17 + nE)))
    if not args.no_e2e:
        e2e = measure_e2e(ctx, pc, src, nE, n_exit, max(1, min(args.steps, 5)))

    subs = {}
    if not args.no_subs and args.workload == "conical" and not args.ids:
        for name in SUB_WORKLOADS:
            try:
                subs[name] = sub_record(ctx, pc, name)
            except Exception as exc:  # a sub-record must not take the headline down
                subs[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    sweep = sweep_c5(ctx, pc, 2.0 if world == 1 else 0.0) if args.sweep else None

    # ---- the library's own multi-device path: ONE process (rank 0) shards one call over all N devices (the other ranks wait)
    one_process = None
    if world > 1 and not args.no_e2e:
        # the idle ranks must not wait in an NCCL barrier: its kernel spins on their GPU and would time-slice with rank 0's kernels there
        cpu_group = dist.new_group(backend="gloo")
        barrier(ctx)
        if rank == 0:
            try:
                one_process = {"what": "polycap_source_get_transmission_efficiencies called ONCE per step by rank 0 with POLYCAP_B200_DEVICES=0,..,N-1: the library "
                                       "shards the photon ids over the %d devices of this run, cuts the id-ordered stream at the n-th exit across them and sums the tallies "
                                       "over peer memory (NVLink); the other ranks idle" % world}
                devs = ",".join(str(i) for i in range(world))  # the N devices of this run, not every device of the box
                one_process[args.workload] = measure_e2e(ctx, pc, src, nE, n_exit * world, max(1, min(args.steps, 5)), devices=devs, all_ranks=False)
                wx = WORKLOADS["xos1"]
                nx = len(wx["energies"])
                one_process["xos1"] = measure_e2e(ctx, pc, build_source(pc, wx), nx, min(200_000 * world, 2_000_000_000 // (8 * (17 + nx))), 3, devices=devs,
                                                  all_ranks=False)
            except Exception as exc:
                one_process = {"error": "%s: %s" % (type(exc).__name__, exc)}
        dist.barrier(group=cpu_group)  # host-side wait (gloo): the other ranks' GPUs stay free for rank 0's call
        barrier(ctx)

    if rank == 0:
        line = {
            "metric": "photons/s", "value": photons / total_s, "unit": "photons/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "reflections_per_s": cnt["refl_all"] / total_s, "exit_photons_per_s": cnt["exit"] / total_s, "photon_ids_per_s": cnt["ids"] / total_s,
            "reflections_of_exit_photons_per_s": cnt["refl_exit"] / total_s,
            "counters": cnt, "efficiencies": (sw / max(photons, 1)).tolist(),
            "gpu_launches": m["launches"], "trace_kernel_launches": m["trace_launches"], "tally_allreduce_ms": m["allreduce_ms"],
            "roofline": roofline(ctx, pc, m, nE, args.workload, kernel_name(nE)),
            "clocks": m["clocks"], "optical_constants": pc.optconst_provider(),
        }
        if e2e:
            line["e2e"] = e2e
        if subs:
            line["sub_records"] = subs
        if one_process:
            line["single_process_multi_gpu"] = one_process
        if sweep:
            line["c5_sweep"] = sweep
        if args.leaks:
            try:
                line["leak_run"] = leak_record(ctx, pc, w, args.leaks, args.cpu_seconds)
            except Exception as exc:
                line["leak_run"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        elif subs and world == 1:
            # BASELINE configs[2] as it is written -- ellipsoidal optic, 400 energies, leak / recap transport ON -- inside the default line
            try:
                line["sub_records"]["ellip_leaks_on"] = leak_record(ctx, pc, WORKLOADS["ellip"], 20000, min(args.cpu_seconds, 10.0))
            except Exception as exc:
                line["sub_records"]["ellip_leaks_on"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                b = reference_sample(w, args.cpu_seconds)
                line["cpu_baseline"] = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["stat_parity"] = stat_parity(line["efficiencies"], photons, b["efficiencies"], b["started"], cnt["exit"], b["exits"])
            except Exception as exc:  # the checker is optional infrastructure; the GPU numbers stand without it
                line["cpu_baseline"] = {"value": None, "unit": "photons/s", "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: %s" % exc}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
