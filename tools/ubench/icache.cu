// Microbenchmark: does a large, desynchronised instruction footprint cap the issue rate (instruction cache)?
// Body of N unfused FP32 instructions per loop iteration (straight-line code of N * 16 bytes), run by W warps per SM
// whose start is staggered so that they sit at different places of the body.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int N> struct Body {
	__device__ __forceinline__ static void run(float (&x)[8], float a, float b) {
		Body<N / 2>::run(x, a, b);
		Body<N - N / 2>::run(x, b, a);
	}
};
template <> struct Body<1> {
	__device__ __forceinline__ static void run(float (&x)[8], float a, float b) {
#pragma unroll
		for (int i = 0; i < 8; i++) x[i] = x[i] * a + b; // 16 instructions
	}
};

template <int N>
__global__ void __launch_bounds__(256) k_body(float* out, float a, float b, int iters, int stagger)
{
	float x[8];
	for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
	// stagger: warp w spins w * stagger dependent additions first
	float s = a;
	for (int i = 0; i < (int)(threadIdx.x >> 5) * stagger + (int)(blockIdx.x % 7) * stagger; i++) s = s * 1.0001f + b;
	for (int it = 0; it < iters; it++)
		Body<N>::run(x, a, b + s * 1e-30f);
	float r = 0; for (int i = 0; i < 8; i++) r += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int N> void bench(float* out, int ctasPerSm)
{
	const int total = 1 << 22; // instructions per thread
	const int iters = total / (N * 16);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int stagger : { 0, 997 })
	{
		float best = 1e9;
		for (int rep = 0; rep < 3; rep++)
		{
			cudaEventRecord(e0);
			k_body<N><<<148 * ctasPerSm, 256>>>(out, 1.0001f, 0.5f, iters, stagger);
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
		}
		CK(cudaGetLastError());
		const double instr = (double)148 * ctasPerSm * 8 * iters * N * 16; // warp instructions
		printf("body %5d instr (%4d KB), %d warps/SM, stagger %4d: %.2f warp-instr/clk/SM\n", N * 16, N * 16 * 16 / 1024, ctasPerSm * 8, stagger, instr / 148 / (best * 1e-3 * 1.965e9));
	}
}

int main()
{
	float* out; CK(cudaMalloc(&out, 148 * 8 * 256 * 4));
	for (int c : { 2, 4 })
	{
		bench<64>(out, c); bench<96>(out, c); bench<128>(out, c); bench<160>(out, c); bench<192>(out, c); bench<256>(out, c);
	}
	return 0;
}
