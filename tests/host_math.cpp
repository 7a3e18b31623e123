// Host build of polycap_b200/csrc/pcb_math.h (the SAME source the CUDA kernels compile) for bit-level checks
// against glibc and against the oracle.  Test infrastructure: built on the fly by tests/test_host_math.py.
#include <complex.h>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "pcb_math.h"

static uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static uint64_t lcg(uint64_t &s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return s >> 11; }
static double uni(uint64_t &s) { return (double)lcg(s) * (1.0 / 9007199254740992.0); }

extern "C" {
// number of inputs (out of n, log-uniform grazing cosines in [lo, hi]) where pcb_acos / pcb_sin / pcb_cos differ from libm
void hm_trig_mismatches(long n, double lo, double hi, long out[3])
{
	uint64_t s = 12345;
	out[0] = out[1] = out[2] = 0;
	for (long i = 0; i < n; i++) {
		double x = lo * exp(uni(s) * log(hi / lo));
		if (i & 1) x = -x;
		double th = acos(x);
		if (bits(pcb_acos(x)) != bits(th)) out[0]++;
		if (bits(pcb_sin(th)) != bits(sin(th))) out[1]++;
		if (bits(pcb_cos(th)) != bits(cos(th))) out[2]++;
	}
}
// small-argument atan/sin/cos (wall taper angle of capil.c:233-240)
void hm_small_trig_mismatches(long n, double lo, double hi, long out[3])
{
	uint64_t s = 999;
	out[0] = out[1] = out[2] = 0;
	for (long i = 0; i < n; i++) {
		double x = lo * exp(uni(s) * log(hi / lo));
		if (i & 1) x = -x;
		double g = atan(x);
		if (bits(pcb_atan(x)) != bits(g)) out[0]++;
		if (bits(pcb_sin(g)) != bits(sin(g))) out[1]++;
		if (bits(pcb_cos(g)) != bits(cos(g))) out[2]++;
	}
}
double hm_acos(double x) { return pcb_acos(x); }
double hm_sin(double x) { return pcb_sin(x); }
double hm_cos(double x) { return pcb_cos(x); }
// complex helpers against C99 complex arithmetic / glibc (what the reference's Fresnel term uses)
void hm_complex_mismatches(long n, long out[3])
{
	uint64_t s = 777;
	out[0] = out[1] = out[2] = 0;
	for (long i = 0; i < n; i++) {
		double re = (uni(s) - 0.3) * pow(10., -6. * uni(s)), im = (uni(s) - 0.5) * pow(10., -8. * uni(s));
		double _Complex z = re + I * im;
		double a, b;
		csqrt_g(re, im, a, b);
		double _Complex r = csqrt(z);
		if (bits(a) != bits(creal(r)) || bits(b) != bits(cimag(r))) out[0]++;
		if (bits(hypot_g(re, im)) != bits(cabs(z))) out[1]++;
		double _Complex q = 1.0 / z;
		cinv_g(re, im, a, b);
		if (bits(a) != bits(creal(q)) || bits(b) != bits(cimag(q))) out[2]++;
	}
}
// one reflection: rtot and the new electric vector (refl_geometry + fresnel_rtot), ec = n_re n_im i2_re i2_im
double hm_refl(const double sn[3], const double dir[3], double ev[3], const double ec[4], double new_ev[3])
{
	V3 n = {sn[0], sn[1], sn[2]}, d = {dir[0], dir[1], dir[2]}, e = {ev[0], ev[1], ev[2]};
	ReflGeom g;
	refl_geometry(n, d, e, g);
	ev[0] = e.x; ev[1] = e.y; ev[2] = e.z;
	new_ev[0] = g.new_ev.x; new_ev[1] = g.new_ev.y; new_ev[2] = g.new_ev.z;
	return fresnel_rtot(ec[0], ec[1], ec[2], ec[3], g.cos_t, g.sin_t, g.frac_s, g.frac_p);
}
// fast vs reference-order Fresnel term on the same (bit-identical) geometry; returns both
void hm_fresnel_pair(const double ec[4], double cos_t, double sin_t, double frac_s, double out[2])
{
	out[0] = fresnel_rtot(ec[0], ec[1], ec[2], ec[3], cos_t, sin_t, frac_s, 1. - frac_s);
	out[1] = fresnel_rtot_fast(ec[0], ec[1], ec[2], ec[3], cos_t, sin_t, frac_s, 1. - frac_s);
}
int hm_segment(const double in[20], double out[6])
{
	V3 c0 = {in[0], in[1], in[2]}, c1 = {in[3], in[4], in[5]}, p0 = {in[8], in[9], in[10]}, dir = {in[14], in[15], in[16]}, hit = {in[17], in[18], in[19]}, sn;
	int r = segment_solve(c0, c1, in[6], in[7], p0, dir, in[19], hit, sn);
	out[0] = hit.x; out[1] = hit.y; out[2] = hit.z; out[3] = sn.x; out[4] = sn.y; out[5] = sn.z;
	return r;
}
}
// pcb_acos_is_obtuse(x) against the test the reference writes, acos(x) > M_PI/2 (capil.c:1270-1273), with glibc's acos:
// a dense sweep of the neighbourhood of the edge (in ulps of the edge) plus log-uniform magnitudes of both signs
extern "C" long hm_obtuse_mismatches(long n)
{
	long bad = 0;
	const double edge = -0x1.cb3b399d747f5p-55;
	double x = edge;
	for (int k = 0; k < 4096; k++) x = nextafter(x, -1.0);
	for (int k = 0; k < 8192; k++, x = nextafter(x, 1.0))
		if (pcb_acos_is_obtuse(x) != (acos(x) > M_PI / 2. || acos(x) < 0.)) bad++;
	uint64_t s = 4242;
	for (long i = 0; i < n; i++) {
		double m = pow(10., -20. * uni(s));
		if (i & 1) m = -m;
		if (pcb_acos_is_obtuse(m) != (acos(m) > M_PI / 2. || acos(m) < 0.)) bad++;
		double f = edge * (0.5 + uni(s)); /* within a factor 2 of the edge */
		if (pcb_acos_is_obtuse(f) != (acos(f) > M_PI / 2. || acos(f) < 0.)) bad++;
	}
	const double sp[] = {0.0, -0.0, 1.0, -1.0, 1.5, -1.5, NAN, 0x1p-55, -0x1p-55, 0x1p-54, -0x1p-54, -0x1p-53, -0x1p-52};
	for (double v : sp)
		if (pcb_acos_is_obtuse(v) != (acos(v) > M_PI / 2. || acos(v) < 0.)) bad++;
	return bad;
}
// pcb_refl_out_of_range / pcb_weight_alive and their high-word screens against the comparisons capil_reflect writes
// (capil.c:636-650): random bit patterns, the neighbourhoods of 0, 1e-4 and 1 in ulps, signed zeros, infinities, NaNs of both signs
extern "C" long hm_predicate_mismatches(long n)
{
	long bad = 0;
	auto chk = [&](double v) {
		if (pcb_refl_out_of_range(v) != (v < 0. || v > 1.)) bad++;
		if (pcb_weight_alive(v) != (v >= 1.e-4)) bad++;
		/* the screens may say "undecided / suspect", never the wrong thing */
		const uint32_t hi = pcb_hi_word(v);
		if (!pcb_refl_suspect_hi(hi) && (v < 0. || v > 1. || v != v)) bad++;
		const int sc = pcb_weight_screen_hi(hi);
		if (sc == 1 && !(v >= 1.e-4)) bad++;
		if (sc == 0 && (v >= 1.e-4)) bad++;
	};
	uint64_t s = 99;
	for (long i = 0; i < n; i++) {
		s = s * 6364136223846793005ULL + 1442695040888963407ULL;
		uint64_t b = s ^ (s >> 29);
		double v;
		memcpy(&v, &b, 8);
		chk(v);
		chk(uni(s));		    /* [0, 1) */
		chk(1.e-4 * (0.5 + uni(s))); /* around the survival threshold */
	}
	const double centres[] = {0.0, 1.e-4, 1.0, -1.e-4, -1.0};
	for (double c : centres) {
		double x = c;
		for (int k = 0; k < 600; k++) x = nextafter(x, -INFINITY);
		for (int k = 0; k < 1200; k++, x = nextafter(x, INFINITY)) chk(x);
	}
	const uint64_t sp[] = {0x0000000000000000ULL, 0x8000000000000000ULL, 0x7ff0000000000000ULL, 0xfff0000000000000ULL, 0x7ff8000000000000ULL,
			       0xfff8000000000000ULL, 0x7fffffffffffffffULL, 0xffffffffffffffffULL, 0x7ff0000000000001ULL, 0x0000000000000001ULL,
			       0x3f1a36e2eb1c432dULL, 0x3f1a36e2eb1c432cULL, 0x3f1a36e2eb1c432eULL, 0x3f1a36e300000000ULL, 0x3f1a36e2ffffffffULL,
			       0x3ff0000000000000ULL, 0x3ff0000000000001ULL, 0x3fefffffffffffffULL, 0x7fefffffffffffffULL};
	for (uint64_t b : sp) {
		double v;
		memcpy(&v, &b, 8);
		chk(v);
	}
	return bad;
}
