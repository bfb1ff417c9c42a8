// sm_100a kernels of the rasterization pipeline. Compiled with --fmad=false: the reference binary
// has no FMA (x86-64 baseline, reference CMakeLists.txt:8), and coverage / depth must match it bit
// for bit, so every multiply and add below rounds separately, in the reference's association.
// IEEE division and square root are nvcc's defaults (-prec-div=true -prec-sqrt=true -ftz=false).
//
//   k_vertex  reference loop A            src/Renderer.cpp:344-345 + htransform :13-20, :195-196
//   k_setup   reference loop C + setup    src/Renderer.cpp:351-380, :163-224, clipTriangle :131-161
//   scanTiles / k_scatter                 16x16 tile binning (no reference counterpart)
//   k_raster  reference loops D/E + shade src/Renderer.cpp:236-305, clear :113-119
#include "mr_types.h"
#include <math.h>

namespace {

struct V3 { float x, y, z; };

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
__global__ void __launch_bounds__(256) k_vertex(const __grid_constant__ FrameParams fp)
{
	const int vi = blockIdx.x * 256 + threadIdx.x;
	const int nTiles = fp.tilesX * fp.tilesY;
	if (vi <= nTiles)
		fp.tileCount[vi] = 0;
	if (vi == 0)
	{
		Counters* c = fp.ctr;
		c->trianglesIn = (unsigned long long)fp.nTriInst; c->records = 0; c->clippedIn = 0; c->pairTotal = 0; c->wideRecords = 0;
		c->overflow = 0; c->ctasDone = 0; c->ovfTotal = 0;
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
// Everything stays in registers (the caller passes scalars by reference and is inlined).
// ------------------------------------------------------------------------------------------
struct Setup
{
	float n1x, n1y, n2x, n2y;
	int x0, x1, y0, y1;
};

__device__ __forceinline__ bool setupTriangle(const FrameParams& fp, const float4 a, const float4 b, const float4 c, Setup& s)
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
	s.n1x = -(a.y - c.y) * i2a;
	s.n1y = (a.x - c.x) * i2a;
	s.n2x = -(b.y - a.y) * i2a;
	s.n2y = (b.x - a.x) * i2a;
	minx = tclamp(minx, 0.0f, w - 1.0f);
	maxx = tclamp(maxx, 0.0f, w - 1.0f);
	miny = tclamp(miny, 0.0f, h - 1.0f);
	maxy = tclamp(maxy, 0.0f, h - 1.0f);
	// Pixel loops: x = floor(minx)+0.5, +1 ... while x <= maxx+0.5 (float sum), same in y.
	// All loop values are exact half-integers, so the last index is floor((max+0.5f) - 0.5f).
	s.x0 = (int)floorf(minx);
	s.y0 = (int)floorf(miny);
	const float xlim = maxx + 0.5f, ylim = maxy + 0.5f;
	int x1 = (int)floorf(xlim - 0.5f), y1 = (int)floorf(ylim - 0.5f);
	if ((float)x1 + 0.5f > xlim) x1--;
	if ((float)(x1 + 1) + 0.5f <= xlim) x1++;
	if ((float)y1 + 0.5f > ylim) y1--;
	if ((float)(y1 + 1) + 0.5f <= ylim) y1++;
	s.x1 = x1;
	s.y1 = y1;
	return true;
}

__device__ __forceinline__ void storeRec(Rec* dst, const float4 a, const float4 b, const float4 c, const Setup& s, int r, int flags, int tri)
{
	float4* d4 = reinterpret_cast<float4*>(dst);
	d4[0] = make_float4(a.x, a.y, c.x, c.y);
	d4[1] = make_float4(s.n1x, s.n1y, s.n2x, s.n2y);
	d4[2] = make_float4(a.w, b.w, c.w, __int_as_float(r));
	d4[3] = make_float4(__uint_as_float((uint32_t)s.x0 | ((uint32_t)s.x1 << 16)), __uint_as_float((uint32_t)s.y0 | ((uint32_t)s.y1 << 16)),
	                    __int_as_float(flags), __int_as_float(tri));
}

// One corner of triangle `tri` of a renderable, in view space (loops A/B/C of paintMesh).
struct Corner
{
	float px, py, pz, nx, ny, nz, u, v;
};

__device__ __forceinline__ Corner fetchCorner(const FrameParams& fp, const MeshDev& m, const RDyn* __restrict__ rd, int tri, int corner)
{
	Corner v;
	const int ip = __ldg(&fp.idxPos[(m.triBase + tri) * 3 + corner]);
	const int in = __ldg(&fp.idxNrm[(m.triBase + tri) * 3 + corner]);
	const float4 p = __ldg(&fp.pos4[m.posBase + ip]);
	const float4 n = __ldg(&fp.nrm4[m.nrmBase + in]);
	const V3 pos = affine(rd->mv, p.x, p.y, p.z);
	const V3 nrm = affine(rd->nm, n.x, n.y, n.z);
	v.px = pos.x; v.py = pos.y; v.pz = pos.z;
	v.nx = nrm.x; v.ny = nrm.y; v.nz = nrm.z;
	v.u = 0.0f; v.v = 0.0f;
	if (m.hasUV)
	{
		const int iu = __ldg(&fp.idxUv[(m.uvTriBase + tri) * 3 + corner]);
		const float2 t = __ldg(&fp.uv2[m.uvBase + iu]);
		v.u = t.x; v.v = t.y;
	}
	return v;
}

// reference clip(), Renderer.cpp:121-129
__device__ __forceinline__ Corner clipEdge(float z, const Corner& a, const Corner& b)
{
	const float k = (fabsf(b.pz - a.pz) < 1e-6f) ? 0.5f : (z - a.pz) / (b.pz - a.pz);
	const float k1 = 1.0f - k;
	Corner v;
	v.px = b.px * k + a.px * k1; v.py = b.py * k + a.py * k1; v.pz = b.pz * k + a.pz * k1;
	v.nx = b.nx * k + a.nx * k1; v.ny = b.ny * k + a.ny * k1; v.nz = b.nz * k + a.nz * k1;
	v.u = b.u * k + a.u * k1;
	v.v = b.v * k + a.v * k1;
	return v;
}

// reference clipTriangle(), Renderer.cpp:131-161, for one requested output triangle `sub`.
// v0..v2 are rotated until v0 has the largest z (at most two rotations), then cut.
// Returns the number of output triangles (1 or 2); o0..o2 receive triangle `sub`.
__device__ __forceinline__ int clipTriangle(float z, Corner v0, Corner v1, Corner v2, int sub, Corner& o0, Corner& o1, Corner& o2)
{
#pragma unroll
	for (int guard = 0; guard < 2; guard++)
		if (v0.pz < v1.pz || v0.pz < v2.pz)
		{
			const Corner t = v0; // swap(v0,v1); swap(v0,v2)  ==  (v0,v1,v2) <- (v2,v0,v1)
			v0 = v2; v2 = v1; v1 = t;
		}
	if (v1.pz > z)
	{
		o0 = clipEdge(z, v0, v2); o1 = clipEdge(z, v1, v2); o2 = v2;
		return 1;
	}
	if (v2.pz > z)
	{
		o0 = clipEdge(z, v0, v1); o1 = v1; o2 = clipEdge(z, v1, v2);
		return 1;
	}
	const Corner v01 = clipEdge(z, v0, v1);
	if (sub == 0)
	{
		o0 = v01; o1 = v1; o2 = v2;
	}
	else
	{
		o0 = v01; o1 = v2; o2 = clipEdge(z, v0, v2);
	}
	return 2;
}

// Emits the (tile, slot, record) pairs of one record with plain per-thread atomics into the
// overflow pair list (slow paths: clipper output, triangles spanning more than MR_SEG_PER_LANE tiles).
__device__ __forceinline__ void emitPairsSerial(const FrameParams& fp, int id, const Setup& s)
{
	const int tyLo = fp.tileRow0, tyHi = fp.tileRow0 + fp.tileRows - 1;
	const int tx0 = s.x0 >> MR_TILE_SHIFT, tx1 = s.x1 >> MR_TILE_SHIFT;
	const int ty0 = max(s.y0 >> MR_TILE_SHIFT, tyLo), ty1 = min(s.y1 >> MR_TILE_SHIFT, tyHi);
	if (ty1 < ty0)
		return;
	const int n = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
	const unsigned long long base = atomicAdd(&fp.ctr->ovfTotal, (unsigned long long)n);
	if (base + (unsigned long long)n > (unsigned long long)fp.pairCap)
	{
		fp.ctr->overflow = 1u;
		return;
	}
	int4* dst = fp.ovfPairs + base;
	for (int ty = ty0; ty <= ty1; ty++)
		for (int tx = tx0; tx <= tx1; tx++)
		{
			const int tile = ty * fp.tilesX + tx;
			*dst++ = make_int4(tile, atomicAdd(&fp.tileCount[tile], 1), id, 0);
		}
}

// Near-plane path of k_setup (rare): rebuilds the corners in view space, clips, sets up, stores
// the records and emits their pairs. Self-contained so that its stack never touches the fast path.
__device__ __noinline__ int setupClipped(const FrameParams& fp, int t, int r, int tri)
{
	const RStat rs = fp.rstat[r];
	const MeshDev m = fp.meshes[rs.mesh];
	const RDyn* rd = &fp.rdyn[r];
	const Corner v0 = fetchCorner(fp, m, rd, tri, 0), v1 = fetchCorner(fp, m, rd, tri, 1), v2 = fetchCorner(fp, m, rd, tri, 2);
	int nrec = 0;
	const int tyLo = fp.tileRow0, tyHi = fp.tileRow0 + fp.tileRows - 1;
	for (int sub = 0; sub < 2; sub++)
	{
		Corner o0, o1, o2;
		const int n = clipTriangle(fp.znear, v0, v1, v2, sub, o0, o1, o2);
		if (sub >= n)
			break;
		const float4 a = project(fp, mk3(o0.px, o0.py, o0.pz)), b = project(fp, mk3(o1.px, o1.py, o1.pz)), c = project(fp, mk3(o2.px, o2.py, o2.pz));
		Setup s;
		if (!setupTriangle(fp, a, b, c, s))
			continue;
		if (min(s.y1 >> MR_TILE_SHIFT, tyHi) < max(s.y0 >> MR_TILE_SHIFT, tyLo))
			continue;
		const int id = 2 * t + sub;
		storeRec(&fp.recs[id], a, b, c, s, r, 1, tri);
		emitPairsSerial(fp, id, s);
		nrec++;
	}
	return nrec;
}

// ------------------------------------------------------------------------------------------
// Kernel 2: near test, clip, setup, and (tile, triangle) pair emission.
// One thread per triangle instance t. A surviving triangle is written at recs[2t] (clipper
// outputs at 2t and 2t+1): the index is the submission id. Its (tile, slot, record) pairs go to
// the warp's private segment of pairs[] (MR_SEG_PER_LANE entries per lane, compacted with a warp
// prefix sum, no global allocation); `slot`, the triangle's rank inside its tile, comes from the
// tile counter with one atomic per distinct tile per warp (__match_any_sync), because
// neighbouring triangles mostly share a tile. Triangles covering more tiles use the overflow
// list. The last CTA to finish turns the tile counters into offsets (exclusive scan).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void scanTiles(const FrameParams& fp, int* sh /* >= 34 ints */)
{
	// exclusive scan of tileCount[0..n) into tileOffset[], by one 256-thread CTA
	const int n = fp.tilesX * fp.tilesY;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int per = (n + 255) / 256;
	const int lo = min(tid * per, n), hi = min(lo + per, n);
	int sum = 0;
	for (int i = lo; i < hi; i++)
		sum += __ldcg(&fp.tileCount[i]);
	int incl = sum;
	for (int o = 1; o < 32; o <<= 1)
	{
		const int u = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o)
			incl += u;
	}
	if (lane == 31)
		sh[wid] = incl;
	__syncthreads();
	if (wid == 0)
	{
		int v = (lane < 8) ? sh[lane] : 0, t = v;
		for (int o = 1; o < 8; o <<= 1)
		{
			const int u = __shfl_up_sync(0xffffffffu, t, o);
			if (lane >= o)
				t += u;
		}
		if (lane < 8)
			sh[lane] = t - v;
		if (lane == 7)
		{
			// total number of (tile, triangle) pairs of the frame; must fit the bin array
			fp.ctr->pairTotal = (unsigned long long)t;
			if (t > fp.pairCap)
				fp.ctr->overflow = 1u;
		}
	}
	__syncthreads();
	int run = sh[wid] + incl - sum;
	for (int i = lo; i < hi; i++)
	{
		fp.tileOffset[i] = run;
		run += __ldcg(&fp.tileCount[i]);
	}
}

__global__ void __launch_bounds__(256) k_setup(const __grid_constant__ FrameParams fp)
{
	__shared__ int sh[48];
	const int t = blockIdx.x * 256 + threadIdx.x;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	bool valid = false;
	int nclip = 0, nrecSlow = 0;
	int tx0 = 0, tx1 = -1, ty0 = 0, ty1 = -1;
	Setup s;
	s.x0 = s.x1 = s.y0 = s.y1 = 0;
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
				nclip = 1;
				nrecSlow = setupClipped(fp, t, r, tri);
			}
		}
		else if (setupTriangle(fp, a, b, c, s))
		{
			tx0 = s.x0 >> MR_TILE_SHIFT;
			tx1 = s.x1 >> MR_TILE_SHIFT;
			ty0 = max(s.y0 >> MR_TILE_SHIFT, fp.tileRow0);
			ty1 = min(s.y1 >> MR_TILE_SHIFT, fp.tileRow0 + fp.tileRows - 1);
			if (ty1 >= ty0)
			{
				valid = true;
				storeRec(&fp.recs[2 * (size_t)t], a, b, c, s, r, 0, tri);
			}
		}
	}
	__syncwarp();

	// ---- pairs: warp-private segment for triangles covering <= MR_SEG_PER_LANE tiles ----
	const int nx = tx1 - tx0 + 1;
	const int ntiles = valid ? nx * (ty1 - ty0 + 1) : 0;
	const bool big = ntiles > MR_SEG_PER_LANE;
	const int npairs = big ? 0 : ntiles;
	int incl = npairs;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const int v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o)
			incl += v;
	}
	const int gw = blockIdx.x * 8 + wid; // global warp index == t / 32
	if (lane == 31)
		fp.warpPairCount[gw] = incl;
	{
		int4* dst = fp.pairs + (size_t)gw * (32 * MR_SEG_PER_LANE) + (incl - npairs);
		const int id = 2 * t;
		const int rounds = __reduce_max_sync(0xffffffffu, npairs);
		for (int k = 0; k < rounds; k++)
		{
			const bool on = k < npairs;
			const int tile = on ? (ty0 + k / nx) * fp.tilesX + tx0 + k % nx : -1 - lane;
			const unsigned peers = __match_any_sync(0xffffffffu, tile);
			const int leader = __ffs(peers) - 1;
			int slot = 0;
			if (on && lane == leader)
				slot = atomicAdd(&fp.tileCount[tile], __popc(peers));
			slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(peers & ((1u << lane) - 1u));
			if (on)
				dst[k] = make_int4(tile, slot, id, 0);
		}
	}
	if (big)
		emitPairsSerial(fp, 2 * t, s);

	// ---- statistics: one atomic per CTA ----
	const int nrecWarp = __reduce_add_sync(0xffffffffu, (valid ? 1 : 0) + nrecSlow);
	const int nclipWarp = __reduce_add_sync(0xffffffffu, nclip);
	if (lane == 0)
	{
		sh[32 + wid] = nrecWarp;
		sh[40 + wid] = nclipWarp;
	}
	// ---- last CTA done: tile counters -> tile offsets ----
	__syncthreads();
	if (threadIdx.x == 0)
	{
		int nr = 0, nc = 0;
		for (int i = 0; i < 8; i++)
		{
			nr += sh[32 + i];
			nc += sh[40 + i];
		}
		if (nr) atomicAdd(&fp.ctr->records, (unsigned long long)nr);
		if (nc) atomicAdd(&fp.ctr->clippedIn, (unsigned long long)nc);
		__threadfence();
		sh[31] = (atomicAdd(&fp.ctr->ctasDone, 1u) == gridDim.x - 1) ? 1 : 0;
	}
	__syncthreads();
	if (sh[31])
	{
		__threadfence();
		scanTiles(fp, sh);
	}
}

// Frames without triangles still need zero offsets for the raster kernel.
__global__ void __launch_bounds__(256) k_scan_only(const __grid_constant__ FrameParams fp)
{
	__shared__ int sh[40];
	scanTiles(fp, sh);
}

// Kernel 3: scatter the pairs into the per-tile bins (a warp per k_setup warp segment, then the
// overflow list).
__global__ void __launch_bounds__(256) k_scatter(const __grid_constant__ FrameParams fp)
{
	if (__ldcg(&fp.ctr->overflow))
		return;
	const int lane = threadIdx.x & 31;
	const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
	const int nWarps = (fp.nTriInst + 31) >> 5;
	if (gw < nWarps)
	{
		const int n = fp.warpPairCount[gw];
		const int4* src = fp.pairs + (size_t)gw * (32 * MR_SEG_PER_LANE);
		for (int k = lane; k < n; k += 32)
		{
			const int4 p = src[k];
			fp.bins[fp.tileOffset[p.x] + p.y] = p.z;
		}
	}
	const unsigned long long ovf = fp.ctr->ovfTotal;
	for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < ovf; i += (unsigned long long)gridDim.x * 256)
	{
		const int4 p = fp.ovfPairs[i];
		fp.bins[fp.tileOffset[p.x] + p.y] = p.z;
	}
}

// ------------------------------------------------------------------------------------------
// Kernel 4: tile rasterizer + shader. One CTA per 16x16 tile, 256 threads.
// Phase 1 (four threads per binned triangle, rows interleaved): reference loops D/E with the
//   float edge chain replayed from the triangle's own bbox start (e += n.x per column,
//   Renderer.cpp:243), depth resolved by 64-bit atomicMin in shared memory on
//   (orderable z) << 32 | (record index + 1). The low word makes equal-z fragments resolve to
//   the earliest submitted triangle, which is what the reference's strict `<` test over in-order
//   submission does.
// Phase 2 (thread per pixel): the winner's barycentrics are re-derived by the same chain, then
//   depth, perspective correction, texture and Blinn-Phong exactly as Renderer.cpp:253-305;
//   pixels without a winner get the clear values (Renderer.cpp:113-119) unless fp.keep.
// ------------------------------------------------------------------------------------------
struct ShadeIn
{
	float k0, k1, k2;
	Corner c0, c1, c2;
};

// Renderer.cpp:271-305 for one pixel; writes image (and the normals image).
__device__ __forceinline__ void shadePixel(const FrameParams& fp, const MatDev& mat, const ShadeIn& in, size_t pix)
{
	const float k0 = in.k0, k1 = in.k1, k2 = in.k2;
	V3 color = mk3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
	V3 value = mk3(mat.emissive[0], mat.emissive[1], mat.emissive[2]);
	if (fp.texturing && mat.texOffset >= 0 && mat.texRows > 0)
	{
		const float u = in.c0.u * k0 + in.c1.u * k1 + in.c2.u * k2;
		const float v = in.c0.v * k0 + in.c1.v * k1 + in.c2.v * k2;
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
		const V3 position = mk3(in.c0.px * k0 + in.c1.px * k1 + in.c2.px * k2, in.c0.py * k0 + in.c1.py * k1 + in.c2.py * k2,
		                        in.c0.pz * k0 + in.c1.pz * k1 + in.c2.pz * k2);
		const V3 light = mk3(fp.light[0], fp.light[1], fp.light[2]);
		const V3 lightdir = fp.lightIsPoint ? normalized3(sub3(light, position)) : light;
		const V3 normal = mk3(in.c0.nx * k0 + in.c1.nx * k1 + in.c2.nx * k2, in.c0.ny * k0 + in.c1.ny * k1 + in.c2.ny * k2,
		                      in.c0.nz * k0 + in.c1.nz * k1 + in.c2.nz * k2);
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
	float* img = fp.image + 3 * pix;
	img[0] = value.x; img[1] = value.y; img[2] = value.z;
}

// Shading of a pixel won by a clipper-made triangle (rare): re-runs the clip to get the corners.
__device__ __noinline__ void shadeClippedPixel(const FrameParams& fp, int r, int tri, int sub, float k0, float k1, float k2, size_t pix)
{
	const RStat rs = fp.rstat[r];
	const MeshDev m = fp.meshes[rs.mesh];
	const RDyn* rd = &fp.rdyn[r];
	const Corner v0 = fetchCorner(fp, m, rd, tri, 0), v1 = fetchCorner(fp, m, rd, tri, 1), v2 = fetchCorner(fp, m, rd, tri, 2);
	ShadeIn in;
	in.k0 = k0; in.k1 = k1; in.k2 = k2;
	clipTriangle(fp.znear, v0, v1, v2, sub, in.c0, in.c1, in.c2);
	const MatDev mat = fp.mats[rd->material];
	shadePixel(fp, mat, in, pix);
}

// Per-warp fragment queue of phase 1. Coverage is sparse (a small triangle covers one or two of
// the ~12 pixel centres of its bbox), so the expensive per-fragment work (exact division, key,
// atomic) is not done inside the divergent scan loop: covered pixels are appended to a queue in
// shared memory and consumed 32 at a time by the whole warp.
#define MR_FQ_CAP 64   // entries per warp (power of two)
#define MR_FQ_SLOTS 16 // triangle slots per warp: 8 quads per iteration, two iterations in flight

struct WarpQueue
{
	float e1[MR_FQ_CAP];
	float e2[MR_FQ_CAP];
	uint32_t info[MR_FQ_CAP]; // pixel index in tile | triangle slot << 8
	float4 tri[MR_FQ_SLOTS];  // d0, d1, d2, record id + 1
};

__device__ __forceinline__ void consumeFragments(const FrameParams& fp, WarpQueue& wq, unsigned long long* keys, int head, int n, int lane)
{
	if (lane < n)
	{
		const int i = (head + lane) & (MR_FQ_CAP - 1);
		const float e1 = wq.e1[i], e2 = wq.e2[i];
		const uint32_t info = wq.info[i];
		const float4 t = wq.tri[info >> 8];
		const float k0 = 1.0f - e1 - e2;
		float z;
		if (fp.persp)
			z = 1.0f / (k0 * t.x + e1 * t.y + e2 * t.z); // Renderer.cpp:255
		else
			z = k0 * t.x + e1 * t.y + e2 * t.z + 0.0f * 1.0f; // Renderer.cpp:261
		if (z == z) // a NaN depth never passes `z < pixdepth`
		{
			const unsigned long long key = ((unsigned long long)zkey(z) << 32) | (unsigned long long)__float_as_uint(t.w);
			unsigned long long* slot = &keys[info & 0xffu];
			if (key < *(volatile unsigned long long*)slot)
				atomicMin(slot, key);
		}
	}
}

__global__ void __launch_bounds__(256) k_raster(const __grid_constant__ FrameParams fp)
{
	__shared__ unsigned long long keys[MR_TILE_PIXELS];
	__shared__ WarpQueue queues[8];
	if (__ldcg(&fp.ctr->overflow))
		return; // the host regrows the pair buffers and re-runs the frame
	const int tx = blockIdx.x;
	const int ty = fp.tileRow0 + blockIdx.y;
	const int tile = ty * fp.tilesX + tx;
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int px = tx * MR_TILE + (tid & 15);
	const int py = ty * MR_TILE + (tid >> 4);
	const bool inImage = px < fp.w && py < fp.h && py >= fp.rowBegin && py < fp.rowEnd;
	const size_t pix = (size_t)py * fp.w + px;
	const int count = fp.tileCount[tile];
	const int tileX0 = tx * MR_TILE, tileY0 = ty * MR_TILE;

	if (count == 0 && !fp.keep)
	{
		// empty tile: clear values only
		if (inImage)
		{
			float* img = fp.image + 3 * pix;
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

	// ---- phase 1: coverage + depth. A quad of lanes per triangle, rows interleaved; all loop
	// bounds are made warp-uniform so that ballots and the queue stay convergent. ----
	{
		WarpQueue& wq = queues[tid >> 5];
		const int* bin = fp.bins + fp.tileOffset[tile];
		const int q = lane & 3;
		const int quad = lane >> 2;
		int qhead = 0, qcount = 0; // warp-uniform
		int parity = 0;
		for (int base = (tid >> 5) * 8; base < count; base += 64, parity ^= 1)
		{
			const int i = base + quad;
			const bool have = i < count;
			float p0x = 0, p0y = 0, p2x = 0, p2y = 0, n1x = 0, n1y = 0, n2x = 0, n2y = 0;
			int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
			const int slot = parity * 8 + quad;
			if (have)
			{
				const int id = __ldg(&bin[i]);
				const float4* r4 = reinterpret_cast<const float4*>(&fp.recs[id]);
				const float4 q0 = __ldg(r4), q1 = __ldg(r4 + 1), q2 = __ldg(r4 + 2), q3 = __ldg(r4 + 3);
				p0x = q0.x; p0y = q0.y; p2x = q0.z; p2y = q0.w;
				n1x = q1.x; n1y = q1.y; n2x = q1.z; n2y = q1.w;
				const uint32_t xspan = __float_as_uint(q3.x), yspan = __float_as_uint(q3.y);
				x0 = xspan & 0xffffu;
				x1 = min((int)(xspan >> 16), tileX0 + MR_TILE - 1);
				y0 = max((int)(yspan & 0xffffu), tileY0);
				y1 = min((int)(yspan >> 16), tileY0 + MR_TILE - 1);
				if (q == 0)
					wq.tri[slot] = make_float4(q2.x, q2.y, q2.z, __uint_as_float((uint32_t)(id + 1)));
			}
			const int xs = max(x0, tileX0);       // first column tested in this tile
			const int ncols = have ? max(x1 - xs + 1, 0) : 0;
			const int myRows = have ? max((y1 - y0 - q + 4) >> 2, 0) : 0; // rows y0+q, y0+q+4, ...
			const float ptx = (float)x0 + 0.5f;
			const int maxRows = __reduce_max_sync(0xffffffffu, myRows);
			const int maxCols = __reduce_max_sync(0xffffffffu, ncols);
			__syncwarp();
			for (int rr = 0; rr < maxRows; rr++)
			{
				const bool rowOn = rr < myRows;
				const int y = y0 + q + 4 * rr;
				const float fy = (float)y + 0.5f;
				float e1 = n1x * (ptx - p2x) + n1y * (fy - p2y);
				float e2 = n2x * (ptx - p0x) + n2y * (fy - p0y);
				if (rowOn)
					for (int x = x0; x < xs; x++) // chain prefix left of the tile
					{
						e1 += n1x;
						e2 += n2x;
					}
				const uint32_t rowInfo = (uint32_t)((y - tileY0) * MR_TILE + (xs - tileX0)) | ((uint32_t)slot << 8);
				for (int cc = 0; cc < maxCols; cc++, e1 += n1x, e2 += n2x)
				{
					const float k0 = 1.0f - e1 - e2;
					const bool inside = rowOn && cc < ncols && !(e1 < 0.0f || e2 < 0.0f || k0 < 0.0f); // Renderer.cpp:245
					const unsigned m = __ballot_sync(0xffffffffu, inside);
					if (m == 0u)
						continue;
					if (inside)
					{
						const int w = (qhead + qcount + __popc(m & ((1u << lane) - 1u))) & (MR_FQ_CAP - 1);
						wq.e1[w] = e1;
						wq.e2[w] = e2;
						wq.info[w] = rowInfo + (uint32_t)cc;
					}
					qcount += __popc(m);
					if (qcount >= 32)
					{
						__syncwarp();
						consumeFragments(fp, wq, keys, qhead, 32, lane);
						qhead = (qhead + 32) & (MR_FQ_CAP - 1);
						qcount -= 32;
					}
				}
			}
			// Triangle slots of this parity are overwritten two iterations from now; fragments
			// still queued then must not refer to them, so drain before reusing a parity.
			if (parity == 1 && qcount > 0)
			{
				__syncwarp();
				consumeFragments(fp, wq, keys, qhead, qcount, lane);
				qhead = (qhead + qcount) & (MR_FQ_CAP - 1);
				qcount = 0;
			}
			__syncwarp();
		}
		if (qcount > 0)
		{
			__syncwarp();
			consumeFragments(fp, wq, keys, qhead, qcount, lane);
		}
	}
	__syncthreads();

	// ---- phase 2: resolve + shade ----
	const uint32_t win = inImage ? (uint32_t)(keys[tid] & 0xffffffffull) : 0u;
	// Replay of the winner's edge chain. Lanes of the same pixel row (a half warp) that share a
	// winner starting left of the tile share the prefix of the chain up to the tile edge: one
	// lane walks it, the others receive it by shuffle and only add their in-tile columns.
	float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
	int id = -1;
	if (win != 0u)
	{
		id = (int)(win - 1u);
		const float4* r4 = reinterpret_cast<const float4*>(&fp.recs[id]);
		q0 = __ldg(r4); q1 = __ldg(r4 + 1); q2 = __ldg(r4 + 2); q3 = __ldg(r4 + 3);
	}
	const int x0 = (int)(__float_as_uint(q3.x) & 0xffffu);
	const float ptx = (float)x0 + 0.5f, fy = (float)py + 0.5f;
	float e1 = q1.x * (ptx - q0.z) + q1.y * (fy - q0.w);
	float e2 = q1.z * (ptx - q0.x) + q1.w * (fy - q0.y);
	int xcur = x0;
	{
		const int prefix = (id >= 0) ? tileX0 - x0 : 0; // columns left of the tile
		const unsigned long long groupKey = (prefix > 0) ? (((unsigned long long)(uint32_t)id << 1) | (unsigned long long)((tid >> 4) & 1))
		                                                 : (0x8000000000000000ull | (unsigned long long)(tid & 31));
		const unsigned peers = __match_any_sync(0xffffffffu, groupKey);
		const int leader = __ffs(peers) - 1;
		if ((tid & 31) == leader && prefix > 0)
			for (int x = 0; x < prefix; x++)
			{
				e1 += q1.x;
				e2 += q1.z;
			}
		const float s1 = __shfl_sync(0xffffffffu, e1, leader), s2 = __shfl_sync(0xffffffffu, e2, leader);
		if (prefix > 0)
		{
			e1 = s1;
			e2 = s2;
			xcur = tileX0;
		}
	}
	if (!inImage)
		return;
	if (win == 0u)
	{
		if (!fp.keep)
		{
			float* img = fp.image + 3 * pix;
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
	for (int x = xcur; x < px; x++)
	{
		e1 += q1.x;
		e2 += q1.z;
	}
	float k0 = 1.0f - e1 - e2, k1 = e1, k2 = e2;
	float z;
	if (fp.persp)
	{
		z = 1.0f / (k0 * q2.x + k1 * q2.y + k2 * q2.z);
		k0 *= q2.x * z;
		k1 *= q2.y * z;
		k2 *= q2.z * z;
	}
	else
		z = k0 * q2.x + k1 * q2.y + k2 * q2.z + 0.0f * 1.0f;
	fp.depth[pix] = z;
	if (fp.winner)
		fp.winner[pix] = id;

	const int r = __float_as_int(q2.w), flags = __float_as_int(q3.z), tri = __float_as_int(q3.w);
	if (flags & 1)
	{
		shadeClippedPixel(fp, r, tri, id & 1, k0, k1, k2, pix);
		return;
	}
	const RDyn* rd = &fp.rdyn[r];
	const MatDev mat = fp.mats[rd->material];
	const bool needGeom = fp.lighting || (fp.texturing && mat.texOffset >= 0 && mat.texRows > 0);
	ShadeIn in;
	in.k0 = k0; in.k1 = k1; in.k2 = k2;
	if (needGeom)
	{
		const RStat rs = fp.rstat[r];
		const MeshDev m = fp.meshes[rs.mesh];
		in.c0 = fetchCorner(fp, m, rd, tri, 0);
		in.c1 = fetchCorner(fp, m, rd, tri, 1);
		in.c2 = fetchCorner(fp, m, rd, tri, 2);
	}
	else
	{
		Corner z0;
		z0.px = z0.py = z0.pz = z0.nx = z0.ny = z0.nz = z0.u = z0.v = 0.0f;
		in.c0 = z0; in.c1 = z0; in.c2 = z0;
	}
	shadePixel(fp, mat, in, pix);
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
	else
		k_scan_only<<<1, 256, 0, stream>>>(fp);
	if (ev) cudaEventRecord(ev[2], stream);
	if (ev) cudaEventRecord(ev[3], stream);
	if (fp.nTriInst > 0)
		k_scatter<<<(fp.nTriInst + 255) / 256, 256, 0, stream>>>(fp);
	if (ev) cudaEventRecord(ev[4], stream);
	if (fp.tileRows > 0)
		k_raster<<<dim3(fp.tilesX, fp.tileRows), 256, 0, stream>>>(fp);
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
