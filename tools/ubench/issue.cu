// Microbenchmarks behind the round-2 kernel design (timing experiments only, not product code):
//   * issue rate of unfused FP32 (FMUL + FADD) against Blackwell's packed FMUL2 / FADD2 (mul/add.rn.f32x2),
//   * FP32 mixed with integer work (do the fma and alu pipes issue side by side?),
//   * cp.async.bulk streaming of ~5 KB blobs into shared memory by persistent CTAs,
//   * 64-bit RED.MIN into an L2-resident key array (the depth test of small triangles),
//   * storing a cleared 1080p framebuffer, launch + event floor.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o issue issue.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

#define ITER 2048
// 8 independent chains of (x = x * a + b) unfused: 16 FP instructions per iteration per thread
__global__ void __launch_bounds__(1024) k_scalar(float* out, float a, float b)
{
	float x[8];
	for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
	for (int it = 0; it < ITER; it++)
#pragma unroll
		for (int i = 0; i < 8; i++) x[i] = x[i] * a + b;
	float s = 0; for (int i = 0; i < 8; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the same 8 chains as 4 packed chains: 8 packed instructions per iteration per thread
__global__ void __launch_bounds__(1024) k_packed(float* out, float a, float b)
{
	u64 x[4];
	for (int i = 0; i < 4; i++) x[i] = pk(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
	const u64 a2 = pk(a, a), b2 = pk(b, b);
	for (int it = 0; it < ITER; it++)
#pragma unroll
		for (int i = 0; i < 4; i++) x[i] = add2(mul2(x[i], a2), b2);
	float s = 0; for (int i = 0; i < 4; i++) { float lo, hi; upk(x[i], lo, hi); s += lo + hi; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 8 packed chains (16 floats): does a deeper packed stream reach 2x the scalar flop rate?
__global__ void __launch_bounds__(1024) k_packed8(float* out, float a, float b)
{
	u64 x[8];
	for (int i = 0; i < 8; i++) x[i] = pk(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
	const u64 a2 = pk(a, a), b2 = pk(b, b);
	for (int it = 0; it < ITER; it++)
#pragma unroll
		for (int i = 0; i < 8; i++) x[i] = add2(mul2(x[i], a2), b2);
	float s = 0; for (int i = 0; i < 8; i++) { float lo, hi; upk(x[i], lo, hi); s += lo + hi; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 4 FP chains + 4 integer chains (LOP3/IADD3 on the alu pipe): 8 FP + 8 INT instructions per iteration
__global__ void __launch_bounds__(1024) k_mixed(float* out, float a, float b, unsigned m)
{
	float x[4]; unsigned y[4];
	for (int i = 0; i < 4; i++) { x[i] = threadIdx.x * 1e-3f + i; y[i] = threadIdx.x + i; }
	for (int it = 0; it < ITER; it++)
#pragma unroll
		for (int i = 0; i < 4; i++) { x[i] = x[i] * a + b; y[i] = (y[i] ^ m) + 0x9e3779b9u; y[i] = (y[i] & m) + it; }
	float s = 0; for (int i = 0; i < 4; i++) s += x[i] + (float)y[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FADD only / FMUL only (which pipe does FADD take?)
__global__ void __launch_bounds__(1024) k_addonly(float* out, float a, float b)
{
	float x[8];
	for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
	for (int it = 0; it < ITER; it++)
#pragma unroll
		for (int i = 0; i < 8; i++) { x[i] = x[i] + a; x[i] = x[i] + b; }
	float s = 0; for (int i = 0; i < 8; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_mulonly(float* out, float a, float b)
{
	float x[8];
	for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i + 1;
	for (int it = 0; it < ITER; it++)
#pragma unroll
		for (int i = 0; i < 8; i++) { x[i] = x[i] * a; x[i] = x[i] * b; }
	float s = 0; for (int i = 0; i < 8; i++) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- cp.async.bulk streaming: CTA of 128 threads, 2 stages of `blob` bytes, n blobs per CTA ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
	asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int STAGES>
__global__ void __launch_bounds__(128) k_bulk(const char* src, int blob, int nblobs, float* out)
{
	extern __shared__ __align__(128) char sm[];
	__shared__ uint64_t full[STAGES];
	if (threadIdx.x == 0) { for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	float acc = 0;
	int issued = 0;
	// blob k of this CTA: global blob index blockIdx.x + k * gridDim.x
	if (threadIdx.x == 0)
		for (; issued < STAGES && issued < nblobs; issued++)
		{
			mbar_expect(&full[issued], blob);
			bulk_g2s(sm + issued * blob, src + (size_t)(blockIdx.x + (size_t)issued * gridDim.x) * blob, blob, &full[issued]);
		}
	for (int k = 0; k < nblobs; k++)
	{
		const int s = k % STAGES;
		mbar_wait(&full[s], (k / STAGES) & 1);
		const float4* p = reinterpret_cast<const float4*>(sm + s * blob);
		for (int i = threadIdx.x; i < blob / 16; i += 128) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
		__syncthreads();
		if (threadIdx.x == 0 && k + STAGES < nblobs)
		{
			mbar_expect(&full[s], blob);
			bulk_g2s(sm + s * blob, src + (size_t)(blockIdx.x + (size_t)(k + STAGES) * gridDim.x) * blob, blob, &full[s]);
		}
	}
	if (acc == 1234.5f) out[0] = acc;
}
// the same bytes with plain coalesced LDG.128 by one-shot CTAs (one blob each)
__global__ void __launch_bounds__(128) k_ldg(const char* src, int blob, float* out)
{
	const float4* p = reinterpret_cast<const float4*>(src + (size_t)blockIdx.x * blob);
	float acc = 0;
	for (int i = threadIdx.x; i < blob / 16; i += 128) { float4 v = __ldg(&p[i]); acc += v.x + v.y + v.z + v.w; }
	if (acc == 1234.5f) out[0] = acc;
}

// ---- 64-bit RED.MIN on a 1080p key array: n atomics per thread at pseudo-random pixels near the thread's home ----
__global__ void __launch_bounds__(128) k_red(u64* keys, int npix, int per)
{
	const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned h = t * 2654435761u;
	const unsigned home = (unsigned)(((u64)t * 7919u) % (unsigned)npix);
	for (int i = 0; i < per; i++)
	{
		h = h * 1664525u + 1013904223u;
		const unsigned p = (home + (h >> 28) + 1920u * ((h >> 24) & 3u)) % (unsigned)npix;
		atomicMin(&keys[p], ((u64)(h | 0x80000000u) << 32) | t);
	}
}

// ---- framebuffer clear by 16x16 tiles, 128 threads, one 32-byte store per thread ----
__global__ void __launch_bounds__(128) k_clear(float* img, float* dep, int w)
{
	const int tid = threadIdx.x, row = tid >> 3, j = tid & 7;
	const size_t pix = (size_t)(blockIdx.y * 16 + row) * w + blockIdx.x * 16;
	float* dst = (j < 6) ? img + 3 * pix + 8 * j : dep + pix + 8 * (j - 6);
	const float v = (j < 6) ? 0.25f : 1e11f;
	asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst), "f"(v) : "memory");
}
__global__ void k_empty() {}
__global__ void k_flush(float4* p, size_t n, float v)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_float4(v, v, v, v);
}

template <class F> float timeit(F f, int reps, void* flushBuf = 0)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	float tot = 0;
	for (int i = 0; i < reps + 2; i++)
	{
		if (flushBuf) k_flush<<<148 * 8, 256>>>((float4*)flushBuf, (size_t)(256 << 20) / 16, (float)i);
		cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		if (i >= 2) tot += ms;
	}
	CK(cudaGetLastError());
	return tot / reps * 1e3f; // us
}

int main()
{
	float* out; CK(cudaMalloc(&out, 148 * 2 * 1024 * 4));
	void* flush; CK(cudaMalloc(&flush, 256 << 20));
	const double thr = 148.0 * 1024 * 2; // two 1024-thread CTAs per SM
	{
		float us;
		us = timeit([&] { k_scalar<<<296, 1024>>>(out, 1.0001f, 0.5f); }, 5);
		printf("scalar  FMUL+FADD   : %7.1f us  %6.2f T thread-instr/s  (%.2f warp-instr/clk/SM at 1.965 GHz)\n", us, thr * ITER * 16 / us * 1e-6, thr * ITER * 16 / 32 / 148 / (us * 1965));
		us = timeit([&] { k_packed<<<296, 1024>>>(out, 1.0001f, 0.5f); }, 5);
		printf("packed4 FMUL2+FADD2 : %7.1f us  %6.2f T thread-instr/s  (%.2f warp-instr/clk/SM) same flops as scalar\n", us, thr * ITER * 8 / us * 1e-6, thr * ITER * 8 / 32 / 148 / (us * 1965));
		us = timeit([&] { k_packed8<<<296, 1024>>>(out, 1.0001f, 0.5f); }, 5);
		printf("packed8 FMUL2+FADD2 : %7.1f us  %6.2f T thread-instr/s  (%.2f warp-instr/clk/SM) 2x the flops of scalar\n", us, thr * ITER * 16 / us * 1e-6, thr * ITER * 16 / 32 / 148 / (us * 1965));
		us = timeit([&] { k_mixed<<<296, 1024>>>(out, 1.0001f, 0.5f, 0x55aa55aau); }, 5);
		printf("mixed   8 FP + ~16 INT: %7.1f us (scalar FP alone would be half of 'scalar')\n", us);
		us = timeit([&] { k_addonly<<<296, 1024>>>(out, 1.0001f, 0.5f); }, 5);
		printf("FADD only           : %7.1f us  (%.2f warp-instr/clk/SM)\n", us, thr * ITER * 16 / 32 / 148 / (us * 1965));
		us = timeit([&] { k_mulonly<<<296, 1024>>>(out, 1.0001f, 0.9999f); }, 5);
		printf("FMUL only           : %7.1f us  (%.2f warp-instr/clk/SM)\n", us, thr * ITER * 16 / 32 / 148 / (us * 1965));
	}
	{
		// 4300 blobs of 4736 bytes (a 130-vertex meshlet) ~ 20 MB; and a larger set
		for (int blob : { 4736, 8192 })
			for (int per : { 29, 116 })
			{
				const int ctasPerSm = 7, grid = 148 * ctasPerSm;
				const int nblobs = (per + ctasPerSm - 1) / ctasPerSm; // per CTA
				const size_t bytes = (size_t)grid * nblobs * blob;
				char* src; CK(cudaMalloc(&src, bytes)); CK(cudaMemset(src, 1, bytes));
				CK(cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192));
				CK(cudaFuncSetAttribute(k_bulk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 8192));
				float us2 = timeit([&] { k_bulk<2><<<grid, 128, 2 * blob>>>(src, blob, nblobs, out); }, 5, flush);
				float us3 = timeit([&] { k_bulk<3><<<grid, 128, 3 * blob>>>(src, blob, nblobs, out); }, 5, flush);
				float usl = timeit([&] { k_ldg<<<grid * nblobs, 128>>>(src, blob, out); }, 5, flush);
				printf("bulk %d B x %d/CTA x %d CTAs (%.1f MB, cold): 2 stages %.1f us (%.0f GB/s), 3 stages %.1f us, one-shot LDG CTAs %.1f us\n", blob, nblobs, grid,
				       bytes * 1e-6, us2, bytes / us2 * 1e-3, us3, usl);
				cudaFree(src);
			}
	}
	{
		const int npix = 1920 * 1080;
		u64* keys; CK(cudaMalloc(&keys, (size_t)npix * 8)); CK(cudaMemset(keys, 0xff, (size_t)npix * 8));
		for (int per : { 1, 2, 4 })
		{
			float us = timeit([&] { k_red<<<(550000 + 127) / 128, 128>>>(keys, npix, per); }, 5);
			printf("RED.MIN.64: 550k threads x %d: %.1f us (%.1f G atomics/s)\n", per, us, 550000.0 * per / us * 1e-3);
		}
		cudaFree(keys);
	}
	{
		float *img, *dep; CK(cudaMalloc(&img, 1920 * 1080 * 12)); CK(cudaMalloc(&dep, 1920 * 1088 * 4));
		float e = timeit([&] { k_empty<<<1, 32>>>(); }, 20);
		float e2 = timeit([&] { k_empty<<<1, 32>>>(); k_empty<<<1, 32>>>(); }, 20);
		float e148 = timeit([&] { k_empty<<<148, 1024>>>(); }, 20);
		float c = timeit([&] { k_clear<<<dim3(120, 67), 128>>>(img, dep, 1920); }, 10, flush);
		float cw = timeit([&] { k_clear<<<dim3(120, 67), 128>>>(img, dep, 1920); }, 10);
		printf("empty kernel %.1f us, two %.1f us, 148x1024 empty %.1f us; clear 33 MB by tiles: %.1f us (L2 flushed), %.1f us (warm)\n", e, e2, e148, c, cw);
	}
	return 0;
}
