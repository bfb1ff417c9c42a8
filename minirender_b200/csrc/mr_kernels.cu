// sm_100a kernels of the rasterization pipeline. Compiled with --fmad=false: the reference binary
// has no FMA (x86-64 baseline, reference CMakeLists.txt:8), and coverage / depth must match it bit
// for bit, so every multiply and add below rounds separately, in the reference's association.
// IEEE division and square root are nvcc's defaults (-prec-div=true -prec-sqrt=true -ftz=false).
//
//   k_vertex  reference loop A            src/Renderer.cpp:344-345 + htransform :13-20, :195-196
//   k_setup   reference loop C + setup    src/Renderer.cpp:351-380, :163-224, clipTriangle :131-161
//   k_scan / k_scatter                    16x16 tile binning (no reference counterpart)
//   k_raster  reference loops D/E + shade src/Renderer.cpp:236-305, clear :113-119
#include "mr_types.h"
#include <math.h>

namespace {

struct V3 { float x, y, z; };
struct Vert { V3 pos; V3 nrm; float u, v; };

__device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 add3(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 scale3(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float len3(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
// asl Vec3::normalized(): multiply by the reciprocal of the length
__device__ __forceinline__ V3 normalized3(V3 a) { float q = 1.0f / len3(a); return mk3(a.x * q, a.y * q, a.z * q); }

// asl ternary min/max/clamp (a NaN compare is false, so the second operand survives)
__device__ __forceinline__ float tmin(float a, float b) { return (a < b) ? a : b; }
__device__ __forceinline__ float tmax(float a, float b) { return (a > b) ? a : b; }
__device__ __forceinline__ float tclamp(float x, float a, float b) { return (x < a) ? a : (x > b) ? b : x; }

// asl::Matrix4 * Vec3 over the top three rows of a row-major 3x4
__device__ __forceinline__ V3 affine(const float* __restrict__ m, float x, float y, float z)
{
	return mk3(m[0] * x + m[1] * y + m[2] * z + m[3],
	           m[4] * x + m[5] * y + m[6] * z + m[7],
	           m[8] * x + m[9] * y + m[10] * z + m[11]);
}

// reference htransform, Renderer.cpp:13-20 (structural zeros are multiplied, not skipped)
__device__ __forceinline__ V3 htransform(const float* m, V3 p)
{
	float iw = 1.0f / (m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15]);
	return mk3((m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3]) * iw,
	           (m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7]) * iw,
	           (m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]) * iw);
}

// view-space vertex -> (pixel x, pixel y, view z, depth term); Renderer.cpp:186-196, :223-224
__device__ __forceinline__ float4 project(const FrameParams& fp, V3 view)
{
	V3 ndc = htransform(fp.P, view);
	float4 o;
	o.x = (1.0f + ndc.x) * (fp.wf / 2.0f);
	o.y = (1.0f - ndc.y) * (fp.hf / 2.0f);
	o.z = view.z;
	o.w = fp.persp ? (-1.0f / view.z) : ndc.z;
	return o;
}

// order-preserving float -> uint map (so that atomicMin on the key is a depth test);
// -0 is folded onto +0 because the reference's `z < pixdepth` treats them as equal
__device__ __forceinline__ uint32_t zkey(float z)
{
	if (z == 0.0f)
		z = 0.0f;
	uint32_t u = __float_as_uint(z);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ int findRenderable(const FrameParams& fp, int start, int inst, bool tri)
{
	int r = start;
	const int last = fp.nRenderables - 1;
	while (r < last)
	{
		const RStat& nx = fp.rstat[r + 1];
		if (inst < (tri ? nx.triBase : nx.vertBase))
			break;
		r++;
	}
	return r;
}

// ------------------------------------------------------------------------------------------
// Kernel 1: vertex transform. One thread per vertex instance: one LDG.128 in, one STG.128 out,
// both fully coalesced. Also zeroes the per-frame tile counters and statistics.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vertex(const FrameParams fp)
{
	const int vi = blockIdx.x * 256 + threadIdx.x;
	const int nTiles = fp.tilesX * fp.tilesY;
	if (vi <= nTiles)
		fp.tileCount[vi] = 0;
	if (vi == 0)
	{
		Counters* c = fp.ctr;
		c->trianglesIn = 0; c->records = 0; c->clippedIn = 0; c->pairTotal = 0; c->wideRecords = 0; c->overflow = 0;
	}
	if (vi >= fp.nVertInst)
		return;
	const int r = findRenderable(fp, fp.vtxBlockR[blockIdx.x], vi, false);
	const RStat rs = fp.rstat[r];
	const MeshDev& m = fp.meshes[rs.mesh];
	const float4 p = __ldg(&fp.pos4[m.posBase + (vi - rs.vertBase)]);
	const V3 view = affine(fp.rdyn[r].mv, p.x, p.y, p.z);
	fp.pv[vi] = project(fp, view);
}

// ------------------------------------------------------------------------------------------
// Triangle setup shared by the direct and the clipped path (Renderer.cpp:198-224).
// a,b,c = projected corners (pixel x, pixel y, view z, depth term). Returns false if rejected.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool setupTriangle(const FrameParams& fp, float4 a, float4 b, float4 c, Rec& rec)
{
	const float w = fp.wf, h = fp.hf;
	float minx = 1e30f, miny = 1e30f, maxx = -1e30f, maxy = -1e30f;
	minx = tmin(minx, a.x); miny = tmin(miny, a.y); maxx = tmax(maxx, a.x); maxy = tmax(maxy, a.y);
	minx = tmin(minx, b.x); miny = tmin(miny, b.y); maxx = tmax(maxx, b.x); maxy = tmax(maxy, b.y);
	minx = tmin(minx, c.x); miny = tmin(miny, c.y); maxx = tmax(maxx, c.x); maxy = tmax(maxy, c.y);
	if (maxx < 0.0f || maxy < 0.0f || minx > w || miny > h)
		return false;
	// (p0 - p1) ^ (p2 - p1)
	const float area = (a.x - b.x) * (c.y - b.y) - (a.y - b.y) * (c.x - b.x);
	// The reference returns on area <= 0. A NaN area gets past that test but then every z it
	// produces is NaN and fails the depth test, so nothing is drawn either.
	if (!(area > 0.0f))
		return false;
	const float i2a = -1.0f / area;
	rec.n1x = -(a.y - c.y) * i2a;
	rec.n1y = (a.x - c.x) * i2a;
	rec.n2x = -(b.y - a.y) * i2a;
	rec.n2y = (b.x - a.x) * i2a;
	minx = tclamp(minx, 0.0f, w - 1.0f);
	maxx = tclamp(maxx, 0.0f, w - 1.0f);
	miny = tclamp(miny, 0.0f, h - 1.0f);
	maxy = tclamp(maxy, 0.0f, h - 1.0f);
	// Pixel loops: x = floor(minx)+0.5, +1 ... while x <= maxx+0.5 (float sum), same in y.
	// All loop values are exact half-integers, so the last index is floor((max+0.5f) - 0.5f).
	const int x0 = (int)floorf(minx), y0 = (int)floorf(miny);
	const float xlim = maxx + 0.5f, ylim = maxy + 0.5f;
	int x1 = (int)floorf(xlim - 0.5f), y1 = (int)floorf(ylim - 0.5f);
	if ((float)x1 + 0.5f > xlim) x1--;
	if ((float)(x1 + 1) + 0.5f <= xlim) x1++;
	if ((float)y1 + 0.5f > ylim) y1--;
	if ((float)(y1 + 1) + 0.5f <= ylim) y1++;
	rec.p0x = a.x; rec.p0y = a.y; rec.p2x = c.x; rec.p2y = c.y;
	rec.d0 = a.w; rec.d1 = b.w; rec.d2 = c.w;
	rec.xspan = (uint32_t)x0 | ((uint32_t)x1 << 16);
	rec.yspan = (uint32_t)y0 | ((uint32_t)y1 << 16);
	return true;
}

// Gathers one corner of triangle `tri` of renderable r in view space (loops A/B/C of paintMesh).
__device__ __forceinline__ Vert fetchCorner(const FrameParams& fp, const MeshDev& m, const RDyn& rd, int tri, int corner)
{
	Vert v;
	const int ip = __ldg(&fp.idxPos[(m.triBase + tri) * 3 + corner]);
	const int in = __ldg(&fp.idxNrm[(m.triBase + tri) * 3 + corner]);
	const float4 p = __ldg(&fp.pos4[m.posBase + ip]);
	const float4 n = __ldg(&fp.nrm4[m.nrmBase + in]);
	v.pos = affine(rd.mv, p.x, p.y, p.z);
	v.nrm = affine(rd.nm, n.x, n.y, n.z);
	if (m.hasUV)
	{
		const int iu = __ldg(&fp.idxUv[(m.uvTriBase + tri) * 3 + corner]);
		const float2 t = __ldg(&fp.uv2[m.uvBase + iu]);
		v.u = t.x; v.v = t.y;
	}
	else
	{
		v.u = 0.0f; v.v = 0.0f;
	}
	return v;
}

// reference clip(), Renderer.cpp:121-129
__device__ __forceinline__ Vert clipEdge(float z, const Vert& a, const Vert& b)
{
	const float k = (fabsf(b.pos.z - a.pos.z) < 1e-6f) ? 0.5f : (z - a.pos.z) / (b.pos.z - a.pos.z);
	const float k1 = 1.0f - k;
	Vert v;
	v.pos = add3(scale3(b.pos, k), scale3(a.pos, k1));
	v.nrm = add3(scale3(b.nrm, k), scale3(a.nrm, k1));
	v.u = b.u * k + a.u * k1;
	v.v = b.v * k + a.v * k1;
	return v;
}

// reference clipTriangle(), Renderer.cpp:131-161. v[] is rotated in place; returns the number of
// output triangles (1 or 2) written to out[k][0..2]. Precondition: not all three beyond z.
__device__ __noinline__ int clipTriangle(float z, Vert* v, Vert (*out)[3])
{
	for (int guard = 0; guard < 3 && (v[0].pos.z < v[1].pos.z || v[0].pos.z < v[2].pos.z); guard++)
	{
		Vert t = v[0]; v[0] = v[1]; v[1] = t;
		t = v[0]; v[0] = v[2]; v[2] = t;
	}
	if (v[1].pos.z > z)
	{
		out[0][0] = clipEdge(z, v[0], v[2]);
		out[0][1] = clipEdge(z, v[1], v[2]);
		out[0][2] = v[2];
		return 1;
	}
	if (v[2].pos.z > z)
	{
		out[0][0] = clipEdge(z, v[0], v[1]);
		out[0][1] = v[1];
		out[0][2] = clipEdge(z, v[1], v[2]);
		return 1;
	}
	const Vert v01 = clipEdge(z, v[0], v[1]);
	const Vert v02 = clipEdge(z, v[0], v[2]);
	out[0][0] = v01; out[0][1] = v[1]; out[0][2] = v[2];
	out[1][0] = v01; out[1][1] = v[2]; out[1][2] = v02;
	return 2;
}

// Near-plane path of k_setup: rebuilds the three corners in view space, clips, sets up.
__device__ __noinline__ int setupClipped(const FrameParams& fp, int r, int tri, Rec* rec)
{
	const RStat rs = fp.rstat[r];
	const MeshDev m = fp.meshes[rs.mesh];
	const RDyn& rd = fp.rdyn[r];
	Vert v[3];
	Vert out[2][3];
	for (int c = 0; c < 3; c++)
		v[c] = fetchCorner(fp, m, rd, tri, c);
	const int n = clipTriangle(fp.znear, v, out);
	int mask = 0;
	for (int k = 0; k < n; k++)
	{
		const float4 a = project(fp, out[k][0].pos), b = project(fp, out[k][1].pos), c = project(fp, out[k][2].pos);
		if (setupTriangle(fp, a, b, c, rec[k]))
			mask |= 1 << k;
	}
	return mask;
}

// ------------------------------------------------------------------------------------------
// Kernel 2: near test, clip, setup, and (tile, triangle) pair emission.
// One thread per triangle instance t. Surviving triangles are written at recs[2t+sub]; their
// (tile, slot, record) pairs are compacted into pairs[] with a warp prefix sum and one
// atomicAdd per warp; `slot` is the triangle's rank inside its tile (from the tile counter).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_setup(const FrameParams fp)
{
	const int t = blockIdx.x * 256 + threadIdx.x;
	const int lane = threadIdx.x & 31;
	Rec rec[2];
	int mask = 0;
	int clipped = 0;
	if (t < fp.nTriInst)
	{
		const int r = findRenderable(fp, fp.triBlockR[blockIdx.x], t, true);
		const RStat rs = fp.rstat[r];
		const MeshDev& m = fp.meshes[rs.mesh];
		const int tri = t - rs.triBase;
		const int* ix = fp.idxPos + (size_t)(m.triBase + tri) * 3;
		const int ia = __ldg(ix), ib = __ldg(ix + 1), ic = __ldg(ix + 2);
		const float4 a = fp.pv[rs.vertBase + ia];
		const float4 b = fp.pv[rs.vertBase + ib];
		const float4 c = fp.pv[rs.vertBase + ic];
		const float zn = fp.znear;
		if (a.z > zn || b.z > zn || c.z > zn) // Renderer.cpp:169-177
		{
			if (!(a.z > zn && b.z > zn && c.z > zn))
			{
				clipped = 1;
				mask = setupClipped(fp, r, tri, rec);
			}
		}
		else if (setupTriangle(fp, a, b, c, rec[0]))
			mask = 1;
		for (int k = 0; k < 2; k++)
			if (mask & (1 << k))
			{
				rec[k].renderable = r;
				rec[k].flags = clipped;
				rec[k].tri = tri;
			}
	}

	// tile ranges, restricted to the tile rows of this frame's strip
	int tx0[2], tx1[2], ty0[2], ty1[2];
	int npairs = 0;
	const int tyLo = fp.tileRow0, tyHi = fp.tileRow0 + fp.tileRows - 1;
	for (int k = 0; k < 2; k++)
	{
		tx0[k] = 0; tx1[k] = -1; ty0[k] = 0; ty1[k] = -1;
		if (mask & (1 << k))
		{
			tx0[k] = (int)(rec[k].xspan & 0xffffu) >> MR_TILE_SHIFT;
			tx1[k] = (int)(rec[k].xspan >> 16) >> MR_TILE_SHIFT;
			ty0[k] = max((int)(rec[k].yspan & 0xffffu) >> MR_TILE_SHIFT, tyLo);
			ty1[k] = min((int)(rec[k].yspan >> 16) >> MR_TILE_SHIFT, tyHi);
			if (ty1[k] >= ty0[k])
			{
				npairs += (tx1[k] - tx0[k] + 1) * (ty1[k] - ty0[k] + 1);
				Rec* dst = &fp.recs[2 * (size_t)t + k];
				const float4* s4 = reinterpret_cast<const float4*>(&rec[k]);
				float4* d4 = reinterpret_cast<float4*>(dst);
				d4[0] = s4[0]; d4[1] = s4[1]; d4[2] = s4[2]; d4[3] = s4[3];
			}
			else
				mask &= ~(1 << k);
		}
	}

	// warp-level compaction of the pair ranges
	int incl = npairs;
	for (int o = 1; o < 32; o <<= 1)
	{
		const int v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o)
			incl += v;
	}
	const int total = __shfl_sync(0xffffffffu, incl, 31);
	const int nrecWarp = __reduce_add_sync(0xffffffffu, __popc(mask));
	const int nclipWarp = __reduce_add_sync(0xffffffffu, clipped);
	const int ninWarp = __reduce_add_sync(0xffffffffu, (t < fp.nTriInst) ? 1 : 0);
	unsigned long long base = 0;
	if (lane == 31)
	{
		if (total > 0)
			base = atomicAdd(&fp.ctr->pairTotal, (unsigned long long)total);
		if (nrecWarp) atomicAdd(&fp.ctr->records, (unsigned long long)nrecWarp);
		if (nclipWarp) atomicAdd(&fp.ctr->clippedIn, (unsigned long long)nclipWarp);
		if (ninWarp) atomicAdd(&fp.ctr->trianglesIn, (unsigned long long)ninWarp);
		if (total > 0 && base + (unsigned long long)total > (unsigned long long)fp.pairCap)
			fp.ctr->overflow = 1u;
	}
	base = __shfl_sync(0xffffffffu, base, 31);
	if (total == 0 || base + (unsigned long long)total > (unsigned long long)fp.pairCap)
		return;
	int4* dst = fp.pairs + base + (incl - npairs);
	for (int k = 0; k < 2; k++)
		if (mask & (1 << k))
		{
			const int id = 2 * t + k;
			for (int ty = ty0[k]; ty <= ty1[k]; ty++)
				for (int tx = tx0[k]; tx <= tx1[k]; tx++)
				{
					const int tile = ty * fp.tilesX + tx;
					const int slot = atomicAdd(&fp.tileCount[tile], 1);
					*dst++ = make_int4(tile, slot, id, 0);
				}
		}
}

// ------------------------------------------------------------------------------------------
// Kernel 3a: exclusive scan of the tile counters (one CTA). 3b: scatter pairs into bins.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_scan(const FrameParams fp)
{
	__shared__ int warpSums[32];
	__shared__ int carry;
	const int n = fp.tilesX * fp.tilesY;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024)
	{
		const int i = base + threadIdx.x;
		const int v = (i < n) ? fp.tileCount[i] : 0;
		int incl = v;
		for (int o = 1; o < 32; o <<= 1)
		{
			const int u = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o)
				incl += u;
		}
		if (lane == 31)
			warpSums[wid] = incl;
		__syncthreads();
		if (wid == 0)
		{
			int s = warpSums[lane];
			for (int o = 1; o < 32; o <<= 1)
			{
				const int u = __shfl_up_sync(0xffffffffu, s, o);
				if (lane >= o)
					s += u;
			}
			warpSums[lane] = s;
		}
		__syncthreads();
		const int prefix = carry + (wid ? warpSums[wid - 1] : 0) + incl - v;
		if (i < n)
			fp.tileOffset[i] = prefix;
		__syncthreads();
		if (threadIdx.x == 1023)
			carry = prefix + v;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) k_scatter(const FrameParams fp)
{
	const Counters* c = fp.ctr;
	if (c->overflow)
		return;
	const unsigned long long total = c->pairTotal;
	for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * 256)
	{
		const int4 p = fp.pairs[i];
		fp.bins[fp.tileOffset[p.x] + p.y] = p.z;
	}
}

// ------------------------------------------------------------------------------------------
// Kernel 4: tile rasterizer + shader. One CTA per 16x16 tile, 256 threads.
// Phase 1 (thread per binned triangle): reference loops D/E with the float edge chain replayed
//   from the triangle's own bbox start (e += n.x per column, Renderer.cpp:243), depth resolved
//   by 64-bit atomicMin in shared memory on (orderable z) << 32 | (record index + 1). The low
//   word makes equal-z fragments resolve to the earliest submitted triangle, which is what the
//   reference's strict `<` test over in-order submission does.
// Phase 2 (thread per pixel): the winner's barycentrics are re-derived by the same chain, then
//   depth, perspective correction, texture and Blinn-Phong exactly as Renderer.cpp:253-305;
//   pixels without a winner get the clear values (Renderer.cpp:113-119) unless fp.keep.
// ------------------------------------------------------------------------------------------
struct TriShade
{
	V3 pos[3];
	V3 nrm[3];
	float u[3], v[3];
};

__device__ __noinline__ void loadTriShade(const FrameParams& fp, const Rec& rec, int sub, TriShade& ts)
{
	const int r = rec.renderable;
	const RStat rs = fp.rstat[r];
	const MeshDev m = fp.meshes[rs.mesh];
	const RDyn& rd = fp.rdyn[r];
	Vert v[3];
	for (int c = 0; c < 3; c++)
		v[c] = fetchCorner(fp, m, rd, rec.tri, c);
	if (rec.flags & 1)
	{
		Vert out[2][3];
		clipTriangle(fp.znear, v, out);
		for (int c = 0; c < 3; c++)
			v[c] = out[sub][c];
	}
	for (int c = 0; c < 3; c++)
	{
		ts.pos[c] = v[c].pos;
		ts.nrm[c] = v[c].nrm;
		ts.u[c] = v[c].u;
		ts.v[c] = v[c].v;
	}
}

__global__ void __launch_bounds__(256) k_raster(const FrameParams fp)
{
	__shared__ unsigned long long keys[MR_TILE_PIXELS];
	const Counters* ctr = fp.ctr;
	if (ctr->overflow)
		return; // the host regrows the pair buffers and re-runs the frame
	const int tx = blockIdx.x % fp.tilesX;
	const int ty = fp.tileRow0 + blockIdx.x / fp.tilesX;
	const int tile = ty * fp.tilesX + tx;
	const int tid = threadIdx.x;
	const int px = tx * MR_TILE + (tid & 15);
	const int py = ty * MR_TILE + (tid >> 4);
	const bool inImage = px < fp.w && py < fp.h && py >= fp.rowBegin && py < fp.rowEnd;
	const size_t pix = (size_t)py * fp.w + px;

	{
		unsigned long long k0 = 0ull; // pixels outside the image / strip can never be won
		if (inImage)
		{
			const float d0 = fp.keep ? fp.depth[pix] : 1e11f;
			k0 = (unsigned long long)zkey(d0) << 32;
		}
		keys[tid] = k0;
	}
	__syncthreads();

	// ---- phase 1: coverage + depth ----
	const int count = fp.tileCount[tile];
	const int* bin = fp.bins + fp.tileOffset[tile];
	const int tileX0 = tx * MR_TILE, tileY0 = ty * MR_TILE;
	for (int i = tid; i < count; i += 256)
	{
		const int id = bin[i];
		const float4* r4 = reinterpret_cast<const float4*>(&fp.recs[id]);
		const float4 q0 = __ldg(r4), q1 = __ldg(r4 + 1), q2 = __ldg(r4 + 2), q3 = __ldg(r4 + 3);
		const float p0x = q0.x, p0y = q0.y, p2x = q0.z, p2y = q0.w;
		const float n1x = q1.x, n1y = q1.y, n2x = q1.z, n2y = q1.w;
		const float d0 = q2.x, d1 = q2.y, d2 = q2.z;
		const uint32_t xspan = __float_as_uint(q3.x), yspan = __float_as_uint(q3.y);
		const int x0 = xspan & 0xffffu, x1 = min((int)(xspan >> 16), tileX0 + MR_TILE - 1);
		const int y0 = max((int)(yspan & 0xffffu), tileY0), y1 = min((int)(yspan >> 16), tileY0 + MR_TILE - 1);
		const float ptx = (float)x0 + 0.5f;
		const unsigned long long low = (unsigned long long)(uint32_t)(id + 1);
		for (int y = y0; y <= y1; y++)
		{
			const float fy = (float)y + 0.5f;
			float e1 = n1x * (ptx - p2x) + n1y * (fy - p2y);
			float e2 = n2x * (ptx - p0x) + n2y * (fy - p0y);
			for (int x = x0; x <= x1; x++, e1 += n1x, e2 += n2x)
			{
				if (x < tileX0)
					continue;
				const float k0 = 1.0f - e1 - e2;
				if (e1 < 0.0f || e2 < 0.0f || k0 < 0.0f)
					continue;
				float z;
				if (fp.persp)
					z = 1.0f / (k0 * d0 + e1 * d1 + e2 * d2);
				else
					z = k0 * d0 + e1 * d1 + e2 * d2 + 0.0f * 1.0f;
				if (!(z == z))
					continue;
				const unsigned long long key = ((unsigned long long)zkey(z) << 32) | low;
				unsigned long long* slot = &keys[(y - tileY0) * MR_TILE + (x - tileX0)];
				if (key < *(volatile unsigned long long*)slot)
					atomicMin(slot, key);
			}
		}
	}
	__syncthreads();

	// ---- phase 2: resolve + shade ----
	if (!inImage)
		return;
	const uint32_t win = (uint32_t)(keys[tid] & 0xffffffffull);
	float* img = fp.image + 3 * pix;
	if (win == 0u)
	{
		if (!fp.keep)
		{
			img[0] = fp.bg[0]; img[1] = fp.bg[1]; img[2] = fp.bg[2];
			fp.depth[pix] = 1e11f;
			if (fp.saveNormals && fp.normals)
			{
				float* pn = fp.normals + 3 * pix;
				pn[0] = 0.0f; pn[1] = 0.0f; pn[2] = 1.0f;
			}
			if (fp.winner)
				fp.winner[pix] = -1;
		}
		return;
	}
	const int id = (int)(win - 1u);
	const Rec rec = fp.recs[id];
	// replay the edge chain of this row up to this pixel
	const int x0 = rec.xspan & 0xffffu;
	const float ptx = (float)x0 + 0.5f, fy = (float)py + 0.5f;
	float e1 = rec.n1x * (ptx - rec.p2x) + rec.n1y * (fy - rec.p2y);
	float e2 = rec.n2x * (ptx - rec.p0x) + rec.n2y * (fy - rec.p0y);
	for (int x = x0; x < px; x++)
	{
		e1 += rec.n1x;
		e2 += rec.n2x;
	}
	float k0 = 1.0f - e1 - e2, k1 = e1, k2 = e2;
	float z;
	if (fp.persp)
	{
		z = 1.0f / (k0 * rec.d0 + k1 * rec.d1 + k2 * rec.d2);
		k0 *= rec.d0 * z;
		k1 *= rec.d1 * z;
		k2 *= rec.d2 * z;
	}
	else
		z = k0 * rec.d0 + k1 * rec.d1 + k2 * rec.d2 + 0.0f * 1.0f;
	fp.depth[pix] = z;
	if (fp.winner)
		fp.winner[pix] = id;

	const MatDev mat = fp.mats[fp.rdyn[rec.renderable].material];
	V3 color = mk3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
	const bool hastexture = fp.texturing && mat.texOffset >= 0 && mat.texRows > 0;
	V3 value = mk3(mat.emissive[0], mat.emissive[1], mat.emissive[2]);
	if (hastexture || fp.lighting)
	{
		TriShade ts;
		loadTriShade(fp, rec, id & 1, ts);
		if (hastexture)
		{
			const float u = ts.u[0] * k0 + ts.u[1] * k1 + ts.u[2] * k2;
			const float v = ts.v[0] * k0 + ts.v[1] * k1 + ts.v[2] * k2;
			const float fv = v - floorf(v), fu = u - floorf(u);
			int ti = (int)(fv * (float)mat.texRows), tj = (int)(fu * (float)mat.texCols);
			// fract() == 1.0f (tiny negative input) indexes one past the end in the reference;
			// clamp instead (documented divergence on UB input)
			ti = min(max(ti, 0), mat.texRows - 1);
			tj = min(max(tj, 0), mat.texCols - 1);
			const float4 tex = __ldg(&fp.texels[mat.texOffset + ti * mat.texCols + tj]);
			color = mk3(tex.x, tex.y, tex.z);
		}
		if (fp.lighting)
		{
			const V3 position = add3(add3(scale3(ts.pos[0], k0), scale3(ts.pos[1], k1)), scale3(ts.pos[2], k2));
			const V3 light = mk3(fp.light[0], fp.light[1], fp.light[2]);
			const V3 lightdir = fp.lightIsPoint ? normalized3(sub3(light, position)) : light;
			const V3 normal = add3(add3(scale3(ts.nrm[0], k0), scale3(ts.nrm[1], k1)), scale3(ts.nrm[2], k2));
			const float nl = dot3(normal, lightdir);
			const float nlen = len3(normal);
			const float d = ((0.0f > nl) ? 0.0f : nl) / nlen + fp.ambient;
			value = add3(value, scale3(color, d));
			if (mat.shininess != 0.0f)
			{
				const V3 viewdir = normalized3(position);
				const V3 hv = sub3(lightdir, viewdir);
				const float hn = dot3(hv, normal);
				const float base = ((hn > 0.0f) ? hn : 0.0f) / (len3(hv) * nlen);
				// the reference's unqualified pow() is the double overload
				const float specular = (float)pow((double)base, (double)mat.shininess);
				value = add3(value, scale3(mk3(mat.specular[0], mat.specular[1], mat.specular[2]), specular));
			}
			if (fp.saveNormals && fp.normals)
			{
				float* pn = fp.normals + 3 * pix;
				pn[0] = normal.x; pn[1] = normal.y; pn[2] = normal.z;
			}
		}
	}
	img[0] = value.x; img[1] = value.y; img[2] = value.z;
}

// ------------------------------------------------------------------------------------------
// Small helpers: range image (Renderer.cpp:388-415), savePPM quantiser (io.cpp:358-361),
// AoS xyz/uv -> padded float4 packing for uploads, FMA-contraction self test.
// ------------------------------------------------------------------------------------------
__global__ void k_range(const float* __restrict__ depth, float* __restrict__ xyz, int w, int h,
                        float p00, float p02, float p11, float p12, float p22, float p23, int persp)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
	if (j >= w || i >= h)
		return;
	const float zfar = persp ? p23 / (p22 + 1.0f) : (p23 - 1.0f) / p22;
	const float fardepth = persp ? zfar : 1.0f;
	const float d = depth[(size_t)i * w + j];
	float* o = xyz + 3 * ((size_t)i * w + j);
	if (d > fardepth)
	{
		o[0] = 0.0f; o[1] = 0.0f; o[2] = 0.0f;
	}
	else
	{
		const float u = ((float)j + 0.5f) / ((float)w / 2.0f) - 1.0f;
		const float v = -((float)i + 0.5f) / ((float)h / 2.0f) + 1.0f;
		const float z = -d;
		o[0] = -(u + p02) * z / p00;
		o[1] = -(v + p12) * z / p11;
		o[2] = z;
	}
}

__global__ void k_rgb8(const float* __restrict__ image, uint8_t* __restrict__ out, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
	{
		float v = image[i] * 255.0f;
		v = (v < 0.0f) ? 0.0f : (v > 255.0f) ? 255.0f : v;
		out[i] = (uint8_t)(int)v; // truncation, like the reference's (byte) cast
	}
}

__global__ void k_pack(float4* __restrict__ dst, const float* __restrict__ src, int n, int comps, float w)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	float4 o;
	o.x = src[(size_t)i * comps];
	o.y = src[(size_t)i * comps + 1];
	o.z = comps > 2 ? src[(size_t)i * comps + 2] : 0.0f;
	o.w = w;
	dst[i] = o;
}

__global__ void k_selftest(const float* in, float* out)
{
	// with contraction, a*b+c keeps the exact product; without, the product rounds first
	out[0] = in[0] * in[1] + in[2];
}

}

void mrk_launch_frame(const FrameParams& fp, cudaStream_t stream, cudaEvent_t* ev)
{
	const int nTiles = fp.tilesX * fp.tilesY;
	const int vthreads = (fp.nVertInst > nTiles + 1) ? fp.nVertInst : nTiles + 1;
	if (ev) cudaEventRecord(ev[0], stream);
	k_vertex<<<(vthreads + 255) / 256, 256, 0, stream>>>(fp);
	if (ev) cudaEventRecord(ev[1], stream);
	if (fp.nTriInst > 0)
		k_setup<<<(fp.nTriInst + 255) / 256, 256, 0, stream>>>(fp);
	if (ev) cudaEventRecord(ev[2], stream);
	k_scan<<<1, 1024, 0, stream>>>(fp);
	if (ev) cudaEventRecord(ev[3], stream);
	if (fp.nTriInst > 0)
		k_scatter<<<148 * 8, 256, 0, stream>>>(fp);
	if (ev) cudaEventRecord(ev[4], stream);
	if (fp.tileRows > 0)
		k_raster<<<fp.tilesX * fp.tileRows, 256, 0, stream>>>(fp);
	if (ev) cudaEventRecord(ev[5], stream);
}

int mrk_selftest_no_fma(cudaStream_t stream)
{
	// a*b is not representable: a = 1+2^-12, b = 1+2^-12 -> exact 1+2^-11+2^-24; c = -(1+2^-11)
	const float h[3] = { 1.0f + 1.0f / 4096.0f, 1.0f + 1.0f / 4096.0f, -(1.0f + 1.0f / 2048.0f) };
	float *d = 0, r = -1.0f;
	if (cudaMalloc(&d, 4 * sizeof(float)) != cudaSuccess)
		return -1;
	cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, stream);
	k_selftest<<<1, 1, 0, stream>>>(d, d + 3);
	cudaMemcpyAsync(&r, d + 3, sizeof(float), cudaMemcpyDeviceToHost, stream);
	cudaError_t e = cudaStreamSynchronize(stream);
	cudaFree(d);
	if (e != cudaSuccess)
		return -1;
	return (r == 0.0f) ? 0 : 1; // fused would give 2^-24
}

void mrk_launch_range(const float* depth, float* xyz, int w, int h, const float* P, cudaStream_t stream)
{
	dim3 grid((w + 255) / 256, h);
	k_range<<<grid, 256, 0, stream>>>(depth, xyz, w, h, P[0], P[2], P[5], P[6], P[10], P[11], P[15] == 0.0f);
}

void mrk_launch_rgb8(const float* image, uint8_t* out, size_t n, cudaStream_t stream)
{
	k_rgb8<<<148 * 8, 256, 0, stream>>>(image, out, n);
}

void mrk_launch_pack(float4* dst4, const float* src, int n, int comps, float w, cudaStream_t stream)
{
	if (n > 0)
		k_pack<<<(n + 255) / 256, 256, 0, stream>>>(dst4, src, n, comps, w);
}
