// Host build of polycap_b200/csrc/pcb_leak_core.cuh (the SAME state machine the CUDA leak kernel compiles, one lane per "warp")
// so that its logic -- frame stack, leak records, and above all the wall march that jumps over its uneventful steps -- can be
// compared on the CPU with the reference's own leak corpus and with the step-by-step march.  Test infrastructure: built on the
// fly by tests/test_host_leak.py; the library itself has no CPU path.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "pcb_leak_core.cuh"

extern "C" void pcb_energy_constants(size_t n_energies, const double *energies, const double *amu, const double *scatf, double density, double *out);

extern "C" {
// replay n start states with leak_calc = true.
// flags: bit 0 = full march everywhere (no uneventful-step shortcut), bit 1 = no jumps (uneventful steps one by one)
// out_state [n][11]: exit coords, exit direction, electric vector, d_travel, i_refl; recs [rec_cap][12+nE]: photon, kind, coords(3), direction(3), elecv(3), n_refl, weights
// work[5]: reflection events, wall-march steps (reference-equivalent), steps evaluated, scan rounds, full solves of the next-bore scan
long hl_replay(int nmax, const double *z, const double *cap, const double *ext, double n_cap, double sig_rough, double density, int nE, const double *energies,
	       const double *amu, const double *scatf, int flags, long n, const double *start /* [n][9] */, int *ret, double *out_state, double *weights,
	       long *leak_n, double *recs, long rec_cap, unsigned long long *work)
{
	const int np = nmax + 1;
	std::vector<double> hexz(np + 1), ec((size_t)EC_ROWS * nE);
	const double n_shells = round(sqrt(12. * n_cap - 3.) / 6. - 0.5);
	for (int i = 0; i < np; i++)
		hexz[i] = ext[i] / (2. * PCB_COSPI_6 * (n_shells + 1));
	hexz[np] = hexz[np - 1];
	pcb_energy_constants((size_t)nE, energies, amu, scatf, density, ec.data());
	std::vector<double> zz(z, z + np), cc(cap, cap + np), ee(ext, ext + np);
	zz.push_back(zz.back()); cc.push_back(cc.back()); ee.push_back(ee.back()); // the reference can read index nmax+1
	LkCtx c{};
	std::vector<double> se(np + 1, 0.), sc(np + 1, 0.), sl(np + 1, 0.), isl(np + 1, 0.);
	for (int i = 0; i <= np; i++) {
		if (i + 1 < np) {
			se[i] = (ee[i + 1] - ee[i]) / (zz[i + 1] - zz[i]);
			sc[i] = (cc[i + 1] - cc[i]) / (zz[i + 1] - zz[i]);
		}
		sl[i] = cc[i] / 10.;
		isl[i] = 1.0 / sl[i];
	}
	c.se = se.data(); c.sc = sc.data(); c.sl = sl.data(); c.isl = isl.data();
	c.nmax = nmax; c.nE = nE; c.mono = (n_shells == 0.); c.monotonic = 1;
	for (int i = 0; i + 1 < np; i++)
		if (!(z[i + 1] > z[i]))
			c.monotonic = 0;
	c.no_fast_march = flags & 1;
	c.no_jump = (flags >> 1) & 1;
	c.n_shells = n_shells; c.sig_rough = sig_rough;
	c.z = zz.data(); c.cap = cc.data(); c.ext = ee.data(); c.hexz = hexz.data(); c.ec = ec.data();
	std::vector<LkFrame> frames(LK_DEPTH);
	std::vector<double> arena((size_t)(LK_DEPTH + 1) * nE), strip((size_t)LK_REC_CAP * (11 + nE));
	long nrec_total = 0;
	LkWork total{};
	for (long i = 0; i < n; i++) {
		const double *s = start + 9 * i;
		V3 sc = {s[0], s[1], s[2]}, sd = {s[3], s[4], s[5]}, sev = {s[6], s[7], s[8]};
		LkOut out{};
		out.rec = strip.data(); out.cap = LK_REC_CAP; out.depth = LK_DEPTH;
		LkFrame *fr = frames.data();
		double *ar = arena.data();
		LkResult r;
		LkWork wk{};
		lk_photon(c, sc, sd, sev, 0.0, fr, ar, out, r, wk);
		if (out.err)
			return -out.err;
		total.refl_events += wk.refl_events; total.wall_steps += wk.wall_steps; total.wall_evals += wk.wall_evals; total.scan_rounds += wk.scan_rounds; total.scan_solves += wk.scan_solves;
		ret[i] = r.ret;
		double *o = out_state + 11 * i;
		o[0] = r.ec.x; o[1] = r.ec.y; o[2] = r.ec.z; o[3] = r.ed.x; o[4] = r.ed.y; o[5] = r.ed.z; o[6] = r.ev.x; o[7] = r.ev.y; o[8] = r.ev.z;
		o[9] = r.dtrav; o[10] = (double)r.irefl;
		memcpy(weights + (size_t)i * nE, ar, (size_t)nE * 8);
		leak_n[i] = out.nrec;
		for (int k = 0; k < out.nrec; k++, nrec_total++) {
			if (nrec_total >= rec_cap)
				return -100;
			double *d = recs + (size_t)nrec_total * (12 + nE);
			d[0] = (double)i;
			memcpy(d + 1, strip.data() + (size_t)k * (11 + nE), (size_t)(11 + nE) * 8);
		}
	}
	work[0] = total.refl_events; work[1] = total.wall_steps; work[2] = total.wall_evals; work[3] = total.scan_rounds; work[4] = total.scan_solves;
	return nrec_total;
}

// polycap_capil_trace_wall on one state; returns the code, fills d_travel, r, q and the march counters
int hl_trace_wall(int nmax, const double *z, const double *cap, const double *ext, double n_cap, int flags, const double *in6, double *d_travel, int *rq, unsigned long long *work)
{
	const int np = nmax + 1;
	const double n_shells = round(sqrt(12. * n_cap - 3.) / 6. - 0.5);
	std::vector<double> zz(z, z + np), cc(cap, cap + np), ee(ext, ext + np), hexz(np + 1);
	zz.push_back(zz.back()); cc.push_back(cc.back()); ee.push_back(ee.back());
	for (int i = 0; i <= np; i++)
		hexz[i] = ee[i] / (2. * PCB_COSPI_6 * (n_shells + 1));
	LkCtx c{};
	std::vector<double> se(np + 1, 0.), sc(np + 1, 0.), sl(np + 1, 0.), isl(np + 1, 0.);
	for (int i = 0; i <= np; i++) {
		if (i + 1 < np) {
			se[i] = (ee[i + 1] - ee[i]) / (zz[i + 1] - zz[i]);
			sc[i] = (cc[i + 1] - cc[i]) / (zz[i + 1] - zz[i]);
		}
		sl[i] = cc[i] / 10.;
		isl[i] = 1.0 / sl[i];
	}
	c.se = se.data(); c.sc = sc.data(); c.sl = sl.data(); c.isl = isl.data();
	c.nmax = nmax; c.nE = 0; c.mono = (n_shells == 0.); c.monotonic = 1;
	c.no_fast_march = flags & 1;
	c.no_jump = (flags >> 1) & 1;
	c.n_shells = n_shells;
	c.z = zz.data(); c.cap = cc.data(); c.ext = ee.data(); c.hexz = hexz.data();
	V3 ec = {in6[0], in6[1], in6[2]}, ed = {in6[3], in6[4], in6[5]};
	LkWork wk{};
	int rc = lk_trace_wall(c, ec, ed, *d_travel, rq[0], rq[1], wk);
	work[0] = wk.wall_steps; work[1] = wk.wall_evals;
	return rc;
}
}
