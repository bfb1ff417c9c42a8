// Microbenchmarks of the memory-streaming floors of the pipeline stages (timing experiments only).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define W 1920
#define H 1080
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// (a) flat resolve: thread = 4 consecutive pixels of a row
__global__ void k_flat(const unsigned long long* __restrict__ keys, float* __restrict__ img, float* __restrict__ dep, int n4)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n4) return;
	const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys) + 2 * (size_t)i;
	ulonglong2 a = k2[0], b = k2[1];
	float v = (a.x & b.x & a.y & b.y) == ~0ull ? 0.25f : 1.0f;
	float4* o = reinterpret_cast<float4*>(img) + 3 * (size_t)i;
	o[0] = make_float4(v, v, v, v); o[1] = make_float4(v, v, v, v); o[2] = make_float4(v, v, v, v);
	reinterpret_cast<float4*>(dep)[i] = make_float4(1e11f, 1e11f, 1e11f, 1e11f);
}
// (a2) flat, stores only
__global__ void k_flat_store(float* __restrict__ img, float* __restrict__ dep, int n4)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n4) return;
	float v = 0.25f;
	float4* o = reinterpret_cast<float4*>(img) + 3 * (size_t)i;
	o[0] = make_float4(v, v, v, v); o[1] = make_float4(v, v, v, v); o[2] = make_float4(v, v, v, v);
	reinterpret_cast<float4*>(dep)[i] = make_float4(1e11f, 1e11f, 1e11f, 1e11f);
}
// (b) tile: CTA of 128 threads per 16x16 tile, like k_raster's empty path
__global__ void __launch_bounds__(128) k_tile(const unsigned long long* __restrict__ keys, float* __restrict__ img, float* __restrict__ dep, int usekeys)
{
	const int tx = blockIdx.x, ty = blockIdx.y, tid = threadIdx.x;
	const int x0 = tx * 16, y0 = ty * 16;
	bool any = false;
	if (usekeys)
		for (int pp = 0; pp < 2; pp++)
		{
			int pi = tid + pp * 128, px = x0 + (pi & 15), py = y0 + (pi >> 4);
			if (py < H) any |= keys[(size_t)py * W + px] != ~0ull;
		}
	float v = __syncthreads_or(any) ? 1.0f : 0.25f;
	for (int pi = tid; pi < 256; pi += 128)
	{
		if (pi < 192)
		{
			int row = pi / 12, j = pi - row * 12, y = y0 + row;
			if (y < H) *reinterpret_cast<float4*>(img + 3 * ((size_t)y * W + x0) + 4 * j) = make_float4(v, v, v, v);
		}
		else
		{
			int t = pi - 192, row = t >> 2, j = t & 3, y = y0 + row;
			if (y < H) *reinterpret_cast<float4*>(dep + (size_t)y * W + x0 + 4 * j) = make_float4(1e11f, 1e11f, 1e11f, 1e11f);
		}
	}
}
// (c) setup floor: idx stream + pv gather + trivial reduce
__global__ void __launch_bounds__(256) k_gather(const int* __restrict__ idx, const float4* __restrict__ pv, int nTri, float* sink, int per)
{
	float acc = 0.f;
	for (int k = 0; k < per; k++)
	{
		int t = (blockIdx.x * per + k) * 256 + threadIdx.x;
		if (t < nTri)
		{
			int ia = idx[3 * (size_t)t], ib = idx[3 * (size_t)t + 1], ic = idx[3 * (size_t)t + 2];
			float4 a = pv[ia], b = pv[ib], c = pv[ic];
			acc += a.x * b.y + c.z;
		}
	}
	if (acc == 1234.5f) *sink = acc;
}
// (c2) same with the loads of `per` triangles batched (ILP)
template <int PER>
__global__ void __launch_bounds__(256) k_gather_ilp(const int* __restrict__ idx, const float4* __restrict__ pv, int nTri, float* sink)
{
	int ia[PER], ib[PER], ic[PER];
	float4 a[PER], b[PER], c[PER];
#pragma unroll
	for (int k = 0; k < PER; k++)
	{
		int t = (blockIdx.x * PER + k) * 256 + threadIdx.x;
		t = min(t, nTri - 1);
		ia[k] = idx[3 * (size_t)t]; ib[k] = idx[3 * (size_t)t + 1]; ic[k] = idx[3 * (size_t)t + 2];
	}
#pragma unroll
	for (int k = 0; k < PER; k++) { a[k] = pv[ia[k]]; b[k] = pv[ib[k]]; c[k] = pv[ic[k]]; }
	float acc = 0.f;
#pragma unroll
	for (int k = 0; k < PER; k++) acc += a[k].x * b[k].y + c[k].z;
	if (acc == 1234.5f) *sink = acc;
}
// (d) vertex floor: float4 in, float4 out
__global__ void __launch_bounds__(256) k_vtx(const float4* __restrict__ in, float4* __restrict__ out, int n)
{
	int i = blockIdx.x * 256 + threadIdx.x;
	if (i < n) { float4 p = in[i]; out[i] = make_float4(p.x * 2.f, p.y + 1.f, p.z, 1.0f / p.x); }
}
__global__ void k_flushread(const float4* src, size_t n, float* sink)
{
	float acc = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = __ldcs(&src[i]); acc += v.x + v.y; }
	if (acc == 12345.678f) *sink = acc;
}

static void* flushBuf; static float* sink;
static void flush(cudaStream_t s)
{
	CK(cudaMemsetAsync(flushBuf, 1, 256u << 20, s));
	k_flushread<<<148 * 8, 256, 0, s>>>((const float4*)((char*)flushBuf + (256u << 20)), (256u << 20) / 16, sink);
}
template <class F> static void timeit(const char* name, F f, double bytes, bool doFlush = true)
{
	cudaStream_t s; CK(cudaStreamCreate(&s));
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e9f, sum = 0; const int reps = 20;
	for (int i = 0; i < reps + 3; i++)
	{
		if (doFlush) flush(s);
		cudaEventRecord(e0, s); f(s); cudaEventRecord(e1, s);
		CK(cudaStreamSynchronize(s));
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (i >= 3) { sum += ms; if (ms < best) best = ms; }
	}
	printf("%-34s avg %7.2f us  best %7.2f us  %7.1f GB/s (avg)\n", name, sum / reps * 1e3, best * 1e3, bytes / (sum / reps * 1e-3) / 1e9);
	cudaStreamDestroy(s);
}

int main()
{
	const int nTri = 1000000, nV = 500002;
	unsigned long long* keys; float *img, *dep; int* idx; float4 *pv, *pos;
	CK(cudaMalloc(&keys, (size_t)W * H * 8)); CK(cudaMalloc(&img, (size_t)W * H * 12)); CK(cudaMalloc(&dep, (size_t)W * H * 4));
	CK(cudaMalloc(&idx, (size_t)nTri * 12)); CK(cudaMalloc(&pv, (size_t)nV * 16)); CK(cudaMalloc(&pos, (size_t)nV * 16));
	CK(cudaMalloc(&flushBuf, (512u << 20) + 256)); CK(cudaMalloc(&sink, 4));
	CK(cudaMemset(keys, 0xff, (size_t)W * H * 8)); CK(cudaMemset(pv, 0, (size_t)nV * 16)); CK(cudaMemset(pos, 0, (size_t)nV * 16));
	std::vector<int> h(3 * (size_t)nTri);
	for (int t = 0; t < nTri; t++) { int q = t / 2, r = q / 1000, c = q % 1000; int a = r * 1000 + c, b = a + 1, d = a + 1000; if (t & 1) { h[3*t] = b; h[3*t+1] = d + 1; h[3*t+2] = d; } else { h[3*t] = a; h[3*t+1] = b; h[3*t+2] = d; } for (int k = 0; k < 3; k++) h[3*t+k] %= nV; }
	CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
	const int n4 = W * H / 4;
	const double fb = (double)W * H * 16, kb = (double)W * H * 8;
	timeit("empty launch", [&](cudaStream_t s) { k_vtx<<<1, 256, 0, s>>>(pos, pv, 1); }, 0);
	timeit("memset 33MB", [&](cudaStream_t s) { cudaMemsetAsync(img, 0, (size_t)W * H * 12, s); cudaMemsetAsync(dep, 0, (size_t)W * H * 4, s); }, fb);
	timeit("flat store-only", [&](cudaStream_t s) { k_flat_store<<<(n4 + 255) / 256, 256, 0, s>>>(img, dep, n4); }, fb);
	timeit("flat keys+store", [&](cudaStream_t s) { k_flat<<<(n4 + 255) / 256, 256, 0, s>>>(keys, img, dep, n4); }, fb + kb);
	timeit("flat keys+store 128thr", [&](cudaStream_t s) { k_flat<<<(n4 + 127) / 128, 128, 0, s>>>(keys, img, dep, n4); }, fb + kb);
	timeit("tile store-only", [&](cudaStream_t s) { k_tile<<<dim3(W / 16, (H + 15) / 16), 128, 0, s>>>(keys, img, dep, 0); }, fb);
	timeit("tile keys+store", [&](cudaStream_t s) { k_tile<<<dim3(W / 16, (H + 15) / 16), 128, 0, s>>>(keys, img, dep, 1); }, fb + kb);
	timeit("tile keys+store (no flush)", [&](cudaStream_t s) { k_tile<<<dim3(W / 16, (H + 15) / 16), 128, 0, s>>>(keys, img, dep, 1); }, fb + kb, false);
	const double gb = (double)nTri * 12 + (double)nV * 16;
	timeit("gather per=1", [&](cudaStream_t s) { k_gather<<<(nTri + 255) / 256, 256, 0, s>>>(idx, pv, nTri, sink, 1); }, gb);
	timeit("gather per=4 serial", [&](cudaStream_t s) { k_gather<<<(nTri + 1023) / 1024, 256, 0, s>>>(idx, pv, nTri, sink, 4); }, gb);
	timeit("gather ilp2", [&](cudaStream_t s) { k_gather_ilp<2><<<(nTri + 511) / 512, 256, 0, s>>>(idx, pv, nTri, sink); }, gb);
	timeit("gather ilp4", [&](cudaStream_t s) { k_gather_ilp<4><<<(nTri + 1023) / 1024, 256, 0, s>>>(idx, pv, nTri, sink); }, gb);
	timeit("gather per=1 (no flush)", [&](cudaStream_t s) { k_gather<<<(nTri + 255) / 256, 256, 0, s>>>(idx, pv, nTri, sink, 1); }, gb, false);
	timeit("vertex 500k", [&](cudaStream_t s) { k_vtx<<<(nV + 255) / 256, 256, 0, s>>>(pos, pv, nV); }, (double)nV * 32);
	timeit("vertex 500k (no flush)", [&](cudaStream_t s) { k_vtx<<<(nV + 255) / 256, 256, 0, s>>>(pos, pv, nV); }, (double)nV * 32, false);
	return 0;
}
