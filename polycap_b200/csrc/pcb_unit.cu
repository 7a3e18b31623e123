/*
 * pcb_unit.cu -- single-item device entry points behind the reference's private test API.
 *
 * The reference's white-box tests call polycap_capil_segment / polycap_refl_polar /
 * polycap_capil_reflect / polycap_capil_trace / polycap_photon_within_pc_boundary directly
 * (src/polycap-private.h:40-49,183-187; tests/capil.c).  Each wrapper here runs the SAME device
 * functions the tracer uses on one item, in a <<<1,1>>> launch, so those goldens exercise the
 * GPU code and not a host re-implementation.  unit_trace is a plain, unvoted, one-photon
 * walk over arbitrary cap_x/cap_y arrays: it doubles as an independent check of the warp state
 * machine in pcb_trace.cuh.
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pcb_device.cuh"
#include "pcb_unit.h"

#define UCK(call)                                     \
	do {                                          \
		cudaError_t e_ = (call);              \
		if (e_ != cudaSuccess) {              \
			g_unit_err = cudaGetErrorString(e_); \
			return -100;                  \
		}                                     \
	} while (0)

static thread_local std::string g_unit_err;
extern "C" const char *pcb_unit_last_error(void) { return g_unit_err.c_str(); }

/* in/out block shared by all unit kernels */
struct UnitIO {
	double d[64];
	int i[8];
};

__global__ void unit_within_kernel(UnitIO *io) { io->i[0] = within_pc(io->d[0], io->d[1], io->d[2]); }

__global__ void unit_segment_kernel(UnitIO *io)
{
	const double *d = io->d;
	V3 c0 = {d[0], d[1], d[2]}, c1 = {d[3], d[4], d[5]};
	double r0 = d[6], r1 = d[7];
	V3 p0 = {d[8], d[9], d[10]}, p1 = {d[11], d[12], d[13]}, dir = {d[14], d[15], d[16]};
	V3 hit = {d[17], d[18], d[19]}, sn = {0, 0, 0};
	int ret;
	/* capil.c:65-100 */
	if (c0.z < 0. || c1.z < 0 || r0 < 0. || r1 < 0. || dir.z < 0 || c0.z != p0.z || c1.z != p1.z || c1.z <= c0.z)
		ret = -1;
	else
		ret = segment_solve(c0, c1, r0, r1, p0, dir, hit.z, hit, sn);
	io->i[0] = ret;
	io->d[20] = hit.x; io->d[21] = hit.y; io->d[22] = hit.z;
	io->d[23] = sn.x; io->d[24] = sn.y; io->d[25] = sn.z;
}

/* d: e, density(unused: constants precomputed), sn(3), dir(3), ev(3); ec: 6 constants */
__global__ void unit_refl_polar_kernel(UnitIO *io)
{
	const double *d = io->d;
	V3 sn = {d[0], d[1], d[2]}, dir = {d[3], d[4], d[5]}, ev = {d[6], d[7], d[8]};
	ReflGeom g;
	refl_geometry(sn, dir, ev, g);
	double rtot = fresnel_rtot(d[9], d[10], d[11], d[12], g.cos_t, g.sin_t, g.frac_s, g.frac_p);
	io->d[20] = rtot;
	io->d[21] = g.new_ev.x; io->d[22] = g.new_ev.y; io->d[23] = g.new_ev.z;
	io->d[24] = ev.x; io->d[25] = ev.y; io->d[26] = ev.z; /* photon->exit_electric_vector after in-place renormalisation */
}

struct UnitPhoton {
	V3 coords, dir, ev;
	double dtrav;
	long long irefl;
};

/* capil_reflect without leaks (capil.c:596-654): returns iesc, updates weights/ev/dir */
__device__ int reflect_plain(const double *ec, int nE, double sig_rough, UnitPhoton &ph, V3 sn, double *w)
{
	norm3(sn);
	norm3(ph.dir);
	double alfa = dot3(ph.dir, sn);
	if (alfa < 0.)
		return -1;
	ReflGeom g;
	refl_geometry(sn, ph.dir, ph.ev, g);
	bool alive = false;
	for (int e = 0; e < nE; e++) {
		double cons1 = ec[EC_KR * nE + e] * alfa * sig_rough;
		double r_rough = exp(-1. * cons1 * cons1);
		double rtot = fresnel_rtot(ec[EC_NRE * nE + e], ec[EC_NIM * nE + e], ec[EC_I2RE * nE + e], ec[EC_I2IM * nE + e], g.cos_t, g.sin_t,
					   g.frac_s, g.frac_p);
		if (rtot < 0. || rtot > 1.)
			return -1;
		w[e] = w[e] * rtot * r_rough;
		if (w[e] >= 1.e-4)
			alive = true;
	}
	ph.ev = g.new_ev;
	return alive ? 1 : 0;
}

__global__ void unit_reflect_kernel(UnitIO *io, const double *ec, int nE, double sig_rough, double *w)
{
	UnitPhoton ph;
	const double *d = io->d;
	ph.dir = {d[0], d[1], d[2]};
	ph.ev = {d[3], d[4], d[5]};
	V3 sn = {d[6], d[7], d[8]};
	io->i[0] = reflect_plain(ec, nE, sig_rough, ph, sn, w);
	io->d[20] = ph.dir.x; io->d[21] = ph.dir.y; io->d[22] = ph.dir.z;
	io->d[23] = ph.ev.x; io->d[24] = ph.ev.y; io->d[25] = ph.ev.z;
}

/* one polycap_capil_trace call (capil.c:1236-1360), plain sequential form */
__global__ void unit_trace_kernel(UnitIO *io, const double *z, const double *cap, const double *ext, const double *cap_x, const double *cap_y,
				  int nmax, int mono, const double *ec, int nE, double sig_rough, double *w)
{
	const double *d = io->d;
	UnitPhoton ph;
	ph.coords = {d[0], d[1], d[2]};
	ph.dir = {d[3], d[4], d[5]};
	ph.ev = {d[6], d[7], d[8]};
	ph.dtrav = d[9];
	ph.irefl = io->i[1];
	int ix = io->i[0];
	int iesc = 0;
	norm3(ph.dir);
	V3 last = ph.coords, pd = ph.dir, sn = {0, 0, 0};
	double cosalfa = 0;
	int ret = 0;
	bool done = false;
	for (int i = ix; i < nmax; i++) {
		V3 c0 = {cap_x[i], cap_y[i], z[i]}, c1 = {cap_x[i + 1], cap_y[i + 1], z[i + 1]};
		V3 p0 = {ph.coords.x + ph.dir.x * (z[i] - ph.coords.z) / ph.dir.z, ph.coords.y + ph.dir.y * (z[i] - ph.coords.z) / ph.dir.z, z[i]};
		if (within_pc(ext[i], c0.x, c0.y) == 0 || within_pc(ext[i + 1], c1.x, c1.y) == 0) {
			ret = -1;
			done = true;
			break;
		}
		V3 hit = last;
		if (c0.z < 0. || c1.z < 0 || cap[i] < 0. || cap[i + 1] < 0. || pd.z < 0 || c1.z <= c0.z) {
			iesc = -1;
			sn = {0, 0, 0};
		} else {
			iesc = segment_solve(c0, c1, cap[i], cap[i + 1], p0, pd, last.z, hit, sn);
		}
		if (iesc == 1)
			last = hit;
		cosalfa = dot3(sn, pd);
		double ac = acos(cosalfa);
		if (ac > PCB_PI / 2. || ac < 0.)
			iesc = -5;
		if (iesc == 1) {
			double ext_here = ((ext[i] - ext[i + 1]) / (z[i] - z[i + 1])) * (last.z - z[i + 1]) + ext[i + 1];
			bool outside = mono ? (sqrt(last.x * last.x + last.y * last.y) >= ext_here) : (within_pc(ext_here, last.x, last.y) == 0);
			if (outside) {
				ret = -3;
				done = true;
			}
			ix = i + 1;
			break;
		} else if (within_pc(ext[i], p0.x, p0.y) == 0) {
			ret = -3;
			done = true;
			break;
		}
	}
	if (!done) {
		if (iesc != 1) {
			ret = -2;
		} else {
			double rx = last.x - ph.coords.x, ry = last.y - ph.coords.y, rz = last.z - ph.coords.z;
			ph.dtrav += sqrt(rx * rx + ry * ry + rz * rz);
			ph.coords = last;
			if (fabs(cosalfa) > 1.0) {
				ret = -1;
			} else {
				for (int i = 0; i < nmax; i++)
					if (z[i] <= ph.coords.z)
						ix = i;
				double ext_here = ((ext[ix + 1] - ext[ix]) / (z[ix + 1] - z[ix])) * (last.z - z[ix]) + ext[ix];
				bool outside = mono ? (ph.coords.x * ph.coords.x + ph.coords.y * ph.coords.y >= ext_here) : (within_pc(ext_here, ph.coords.x, ph.coords.y) == 0);
				if (outside) {
					ret = -3;
				} else {
					ret = reflect_plain(ec, nE, sig_rough, ph, sn, w);
					if (ret == 1) {
						ph.dir.x = ph.dir.x - 2.0 * cosalfa * sn.x;
						ph.dir.y = ph.dir.y - 2.0 * cosalfa * sn.y;
						ph.dir.z = ph.dir.z - 2.0 * cosalfa * sn.z;
						norm3(ph.dir);
						ph.irefl++;
					} else if (ret == -1 || ret == -2) {
						ret = -1;
					}
				}
			}
		}
	}
	io->i[0] = ix;
	io->i[1] = (int)ph.irefl;
	io->i[2] = ret;
	io->d[20] = ph.coords.x; io->d[21] = ph.coords.y; io->d[22] = ph.coords.z;
	io->d[23] = ph.dir.x; io->d[24] = ph.dir.y; io->d[25] = ph.dir.z;
	io->d[26] = ph.ev.x; io->d[27] = ph.ev.y; io->d[28] = ph.ev.z;
	io->d[29] = ph.dtrav;
}

/* source.c:23-144 driven by a host-drawn buffer of uniforms (the polycap_rng stream) */
struct BufferRng {
	const double *u;
	int n, used;
	bool overrun;
	__device__ double uniform()
	{
		if (used >= n) {
			overrun = true;
			return 0.5;
		}
		return u[used++];
	}
};

__global__ void unit_get_photon_kernel(UnitIO *io, PcbSource s, int mono, const double *u, int nu)
{
	BufferRng rng = {u, nu, 0, false};
	double r = rng.uniform();
	double phi = atan(s.src_y / s.src_x * tan(2.0 * PCB_PI * r / 4.));
	r = rng.uniform();
	if ((r >= 0.25) && (r < 0.5)) phi = PCB_PI - phi;
	if ((r >= 0.5) && (r < 0.75)) phi = PCB_PI + phi;
	if (r >= 0.75) phi = -1.0 * phi;
	double cphi = cos(phi), sphi = sin(phi);
	double max_rad = s.src_x * s.src_y / sqrt((s.src_y * cphi) * (s.src_y * cphi) + (s.src_x * sphi) * (s.src_x * sphi));
	r = rng.uniform();
	double sx = sqrt(r) * max_rad * cphi + s.src_shiftx;
	double sy = sqrt(r) * max_rad * sphi + s.src_shifty;
	V3 c, dir, ev;
	if (s.src_sigx < 0. || s.src_sigy < 0.) {
		if (mono) {
			r = rng.uniform();
			c.x = (2. * r - 1.) * s.cap0;
			r = rng.uniform();
			c.y = (2. * r - 1.) * s.cap0;
		} else {
			int inside;
			do {
				r = rng.uniform();
				c.x = (2. * r - 1.) * s.ext0;
				r = rng.uniform();
				c.y = (2. * r - 1.) * s.ext0;
				inside = within_pc(s.ext0, c.x, c.y);
			} while (inside == 0 && !rng.overrun);
		}
		dir.x = c.x - sx; dir.y = c.y - sy; dir.z = s.d_source;
	} else {
		r = rng.uniform();
		dir.x = s.src_sigx * (1. - 2. * fabs(r));
		r = rng.uniform();
		dir.y = s.src_sigy * (1. - 2. * fabs(r));
		dir.z = 1.;
		c.x = sx + dir.x * s.d_source / dir.z;
		c.y = sy + dir.y * s.d_source / dir.z;
	}
	c.z = 0.;
	norm3(dir);
	double frac_hor = (1. + s.hor_pol) / 2.;
	r = rng.uniform();
	if (fabs(r) <= frac_hor) ev = {1., 0., 0.};
	else ev = {0., 1., 0.};
	double cosalpha = dot3(ev, dir);
	double alpha = acos(cosalpha);
	double c_ae = 1. / sin(alpha);
	double c_be = -1. * c_ae * cosalpha;
	ev.x = ev.x * c_ae + dir.x * c_be;
	ev.y = ev.y * c_ae + dir.y * c_be;
	ev.z = ev.z * c_ae + dir.z * c_be;
	norm3(ev);
	io->d[0] = c.x; io->d[1] = c.y; io->d[2] = c.z;
	io->d[3] = dir.x; io->d[4] = dir.y; io->d[5] = dir.z;
	io->d[6] = ev.x; io->d[7] = ev.y; io->d[8] = ev.z;
	io->d[9] = sx; io->d[10] = sy;
	io->i[0] = rng.used;
	io->i[1] = rng.overrun ? 1 : 0;
}

/* ------------------------------------------------------------------ host wrappers */

struct DevBuf {
	void *p = nullptr;
	~DevBuf() { cudaFree(p); }
	cudaError_t alloc(size_t n) { return cudaMalloc(&p, n); }
};

static int run_io(UnitIO &io, void (*launch)(UnitIO *, void *), void *ctx)
{
	DevBuf b;
	UCK(b.alloc(sizeof(UnitIO)));
	UCK(cudaMemcpy(b.p, &io, sizeof io, cudaMemcpyHostToDevice));
	launch((UnitIO *)b.p, ctx);
	UCK(cudaGetLastError());
	UCK(cudaMemcpy(&io, b.p, sizeof io, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int pcb_unit_within(double R, double x, double y)
{
	UnitIO io{};
	io.d[0] = R; io.d[1] = x; io.d[2] = y;
	int rc = run_io(io, [](UnitIO *d, void *) { unit_within_kernel<<<1, 1>>>(d); }, nullptr);
	return rc ? rc : io.i[0];
}

extern "C" int pcb_unit_segment(const double in[20], double out[6])
{
	UnitIO io{};
	memcpy(io.d, in, 20 * 8);
	int rc = run_io(io, [](UnitIO *d, void *) { unit_segment_kernel<<<1, 1>>>(d); }, nullptr);
	if (rc)
		return rc;
	memcpy(out, io.d + 20, 6 * 8);
	return io.i[0];
}

extern "C" int pcb_unit_refl_polar(const double in[13], double out[7])
{
	UnitIO io{};
	memcpy(io.d, in, 13 * 8);
	int rc = run_io(io, [](UnitIO *d, void *) { unit_refl_polar_kernel<<<1, 1>>>(d); }, nullptr);
	if (rc)
		return rc;
	memcpy(out, io.d + 20, 7 * 8);
	return 0;
}

extern "C" int pcb_unit_reflect(const double in[9], const double *ec, int nE, double sig_rough, double *w, double out[6])
{
	UnitIO io{};
	memcpy(io.d, in, 9 * 8);
	DevBuf dec, dw, dio;
	UCK(dec.alloc((size_t)EC_ROWS * nE * 8));
	UCK(dw.alloc((size_t)nE * 8));
	UCK(dio.alloc(sizeof io));
	UCK(cudaMemcpy(dec.p, ec, (size_t)EC_ROWS * nE * 8, cudaMemcpyHostToDevice));
	UCK(cudaMemcpy(dw.p, w, (size_t)nE * 8, cudaMemcpyHostToDevice));
	UCK(cudaMemcpy(dio.p, &io, sizeof io, cudaMemcpyHostToDevice));
	unit_reflect_kernel<<<1, 1>>>((UnitIO *)dio.p, (const double *)dec.p, nE, sig_rough, (double *)dw.p);
	UCK(cudaGetLastError());
	UCK(cudaMemcpy(&io, dio.p, sizeof io, cudaMemcpyDeviceToHost));
	UCK(cudaMemcpy(w, dw.p, (size_t)nE * 8, cudaMemcpyDeviceToHost));
	memcpy(out, io.d + 20, 6 * 8);
	return io.i[0];
}

extern "C" int pcb_unit_trace(const double in[10], int *ix, long long *irefl, const double *z, const double *cap, const double *ext,
			      const double *cap_x, const double *cap_y, int nmax, int mono, const double *ec, int nE, double sig_rough, double *w,
			      double out[10])
{
	UnitIO io{};
	memcpy(io.d, in, 10 * 8);
	io.i[0] = *ix;
	io.i[1] = (int)*irefl;
	const size_t tb = (size_t)(nmax + 2) * 8;
	DevBuf dt, dec, dw, dio;
	UCK(dt.alloc(5 * tb));
	UCK(dec.alloc((size_t)EC_ROWS * nE * 8));
	UCK(dw.alloc((size_t)nE * 8));
	UCK(dio.alloc(sizeof io));
	UCK(cudaMemset(dt.p, 0, 5 * tb));
	const double *src[5] = {z, cap, ext, cap_x, cap_y};
	for (int k = 0; k < 5; k++)
		UCK(cudaMemcpy((char *)dt.p + k * tb, src[k], (size_t)(nmax + 1) * 8, cudaMemcpyHostToDevice));
	UCK(cudaMemcpy(dec.p, ec, (size_t)EC_ROWS * nE * 8, cudaMemcpyHostToDevice));
	UCK(cudaMemcpy(dw.p, w, (size_t)nE * 8, cudaMemcpyHostToDevice));
	UCK(cudaMemcpy(dio.p, &io, sizeof io, cudaMemcpyHostToDevice));
	const double *T = (const double *)dt.p;
	const size_t tn = tb / 8;
	unit_trace_kernel<<<1, 1>>>((UnitIO *)dio.p, T, T + tn, T + 2 * tn, T + 3 * tn, T + 4 * tn, nmax, mono, (const double *)dec.p, nE, sig_rough,
				    (double *)dw.p);
	UCK(cudaGetLastError());
	UCK(cudaMemcpy(&io, dio.p, sizeof io, cudaMemcpyDeviceToHost));
	UCK(cudaMemcpy(w, dw.p, (size_t)nE * 8, cudaMemcpyDeviceToHost));
	*ix = io.i[0];
	*irefl = io.i[1];
	memcpy(out, io.d + 20, 10 * 8);
	return io.i[2];
}

extern "C" int pcb_unit_get_photon(const double src[10], int mono, const double *u, int nu, double out[11], int *used)
{
	PcbSource s{};
	s.d_source = src[0]; s.src_x = src[1]; s.src_y = src[2]; s.src_sigx = src[3]; s.src_sigy = src[4];
	s.src_shiftx = src[5]; s.src_shifty = src[6]; s.hor_pol = src[7]; s.cap0 = src[8]; s.ext0 = src[9];
	UnitIO io{};
	DevBuf du, dio;
	UCK(du.alloc((size_t)nu * 8));
	UCK(dio.alloc(sizeof io));
	UCK(cudaMemcpy(du.p, u, (size_t)nu * 8, cudaMemcpyHostToDevice));
	unit_get_photon_kernel<<<1, 1>>>((UnitIO *)dio.p, s, mono, (const double *)du.p, nu);
	UCK(cudaGetLastError());
	UCK(cudaMemcpy(&io, dio.p, sizeof io, cudaMemcpyDeviceToHost));
	memcpy(out, io.d, 11 * 8);
	*used = io.i[0];
	return io.i[1] ? -1 : 0;
}
