// FP64 latency / throughput microbenchmark for sm_100a (tuning aid, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_latency fp64_latency.cu
// Prints the dependent-issue latency of DFMA, IEEE division, sqrt and the reciprocal seed, and DFMA throughput per SM
// as a function of resident warps with ONE dependent chain per thread (what a photon's trajectory arithmetic looks like).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(double *out, long long *cyc, int iters, int what)
{
	double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		if (what == 0) { a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); }
		if (what == 1) { a = a / b; a = a / b; a = a / b; a = a / b; }
		if (what == 2) { a = sqrt(a) + 1.0; a = sqrt(a) + 1.0; a = sqrt(a) + 1.0; a = sqrt(a) + 1.0; }
		if (what == 3) { a = a * b; a = a * b; a = a * b; a = a * b; }
		if (what == 4) { a = a + c; a = a + c; a = a + c; a = a + c; }
	}
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = a;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		*cyc = t1 - t0;
}

__global__ void thr_kernel(double *out, int iters)
{
	double a = 1.0 + threadIdx.x * 1e-9;
	const double b = 1.0000001, c = 1e-9;
	for (int i = 0; i < iters; i++) {
		a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c);
		a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c); a = __fma_rn(a, b, c);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

__global__ void thr2_kernel(double *out, int iters)
{
	double a = 1.0 + threadIdx.x * 1e-9, a2 = 2.0 + threadIdx.x * 1e-9;
	const double b = 1.0000001, c = 1e-9;
	for (int i = 0; i < iters; i++) {
		a = __fma_rn(a, b, c); a2 = __fma_rn(a2, b, c); a = __fma_rn(a, b, c); a2 = __fma_rn(a2, b, c);
		a = __fma_rn(a, b, c); a2 = __fma_rn(a2, b, c); a = __fma_rn(a, b, c); a2 = __fma_rn(a2, b, c);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = a + a2;
}

int main()
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	const int sms = prop.multiProcessorCount;
	const double ghz = prop.clockRate * 1e-6;
	double *out;
	long long *cyc, h;
	cudaMalloc(&out, (size_t)sms * 2048 * 8);
	cudaMalloc(&cyc, 8);
	const char *names[] = {"DFMA", "IEEE div", "sqrt(+DADD)", "DMUL", "DADD"};
	for (int what = 0; what < 5; what++) {
		lat_kernel<<<1, 32>>>(out, cyc, 1000, what);
		lat_kernel<<<1, 32>>>(out, cyc, 20000, what);
		cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
		printf("latency %-12s %.1f cycles per dependent op (one warp alone)\n", names[what], (double)h / (20000.0 * 4));
	}
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int pass = 0; pass < 2; pass++)
		for (int warps = 4; warps <= 64; warps += (warps < 32 ? 4 : 8)) {
			const int iters = 20000;
			const int threads = warps * 32 > 1024 ? warps * 16 : warps * 32, blocks = sms * (warps * 32 / threads);
			if (pass == 0) thr_kernel<<<blocks, threads>>>(out, 100); else thr2_kernel<<<blocks, threads>>>(out, 100);
			cudaEventRecord(e0);
			if (pass == 0) thr_kernel<<<blocks, threads>>>(out, iters); else thr2_kernel<<<blocks, threads>>>(out, iters);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms;
			cudaEventElapsedTime(&ms, e0, e1);
			const double dfma = 8.0 * iters * (double)blocks * threads;
			printf("%d chain(s)/thread, %2d warps/SM: %.2f TFLOP/s, %.1f DFMA/clk/SM (nominal %.2f GHz)\n", pass + 1, warps, 2 * dfma / (ms * 1e-3) / 1e12,
			       dfma / (ms * 1e-3) / sms / (ghz * 1e9), ghz);
		}
	return 0;
}
