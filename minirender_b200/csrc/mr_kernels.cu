// sm_100a kernels of the rasterization pipeline. Compiled with --fmad=false: the reference binary
// has no FMA (x86-64 baseline, reference CMakeLists.txt:8), and coverage / depth must match it bit
// for bit, so every multiply and add below rounds separately, in the reference's association.
// IEEE division and square root are nvcc's defaults (-prec-div=true -prec-sqrt=true -ftz=false).
//
//   k_geom     reference loops A, B, C     src/Renderer.cpp:344-380 + htransform :13-20, :195-196,
//              setup + clipping            :163-224, clipTriangle :131-161,
//              loops D/E of small triangles :238-269 (depth test = 64-bit atomicMin per covered pixel)
//   k_raster   loops D/E of the larger triangles, resolve + shade src/Renderer.cpp:236-305, clear :113-119
#include "mr_types.h"
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-
// serialization attribute may start while its predecessor drains; it must not touch the
// predecessor's output before pdlWait(). pdlLaunchDependents() in the predecessor lets the
// successor's CTAs take the SM slots that free up during the predecessor's last wave.
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdlLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;"); }

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

// The same for the standard perspective matrix (fp.stdProj: rows (a 0 b 0) (0 c d 0) (. . . .) (0 0 -1 0), near plane in
// front of the eye) and finite x, y: the structural zeros contribute +-0 terms only, which change nothing but the sign
// of a zero sum, and the trailing `+ 0.0f` (the reference's + m[3]) settles that the same way; w = -z makes the depth
// term -1/z the very reciprocal htransform computes (IEEE division is sign-symmetric), except at z = +-0, where such a
// corner is in front of the near plane and the triangle takes the clip path, which projects its own corners.
__device__ __forceinline__ float4 projectStd(const FrameParams& fp, V3 view)
{
	const float iw = 1.0f / (0.0f - view.z);
	const float nx = (fp.P[0] * view.x + fp.P[2] * view.z + 0.0f) * iw;
	const float ny = (fp.P[5] * view.y + fp.P[6] * view.z + 0.0f) * iw;
	float4 o;
	o.x = (1.0f + nx) * (fp.wf / 2.0f);
	o.y = (1.0f - ny) * (fp.hf / 2.0f);
	o.z = view.z;
	o.w = iw;
	return o;
}

// order-preserving float -> uint map (so that atomicMin on the key is a depth test);
// -0 is folded onto +0 because the reference's `z < pixdepth` treats them as equal
__device__ __forceinline__ float unzkey(uint32_t k) // the float a depth key was made of (+0 for either zero)
{
	return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ uint32_t zkey(float z)
{
	if (z == 0.0f)
		z = 0.0f;
	uint32_t u = __float_as_uint(z);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Per-frame tables: in the kernel parameters for small scenes, in global memory otherwise.
// The kernels are compiled once per table mode TM, so that every table access is a plain global
// load (TM_GLOBAL) or a constant-bank operand of the kernel parameters (TM_INLINE) instead of a
// generic-address load through a run-time pointer select.
enum { TM_GLOBAL = 0, TM_INLINE = 1 };
template <int TM> __device__ __forceinline__ const RDyn* frameRdyn(const FrameParams& fp) { return TM == TM_INLINE ? fp.rdynInline : fp.rdyn; }
template <int TM> __device__ __forceinline__ const MatDev* frameMats(const FrameParams& fp) { return TM == TM_INLINE ? fp.matsInline : fp.mats; }
template <int TM> __device__ __forceinline__ const RStat* frameRstat(const FrameParams& fp) { return TM == TM_INLINE ? fp.rstatInline : fp.rstat; }

// Cluster culling (k_geom's first phase, one thread per cluster of the frame). Cluster `ci` = MR_CLUSTER
// consecutive triangle instances of one renderable = MR_CLUSTER consecutive triangles of its mesh,
// for which mr_upload_scene stored a bounding sphere and a cone around the geometric normals.
// Returns false when no triangle of the cluster can reach a pixel of this frame:
//  * every vertex nearer than the near plane: the reference drops such triangles (Renderer.cpp:169-177);
//  * the sphere outside the columns / rows the frame can touch (widened by 1.5 px): off-screen reject
//    (Renderer.cpp:202) or, for strip rendering, another rank's rows;
//  * every normal of the cone facing away from the eye: area cull (Renderer.cpp:205-210).
// Only triangles the reference discards itself are dropped, with generous margins, so the image
// does not change.
__device__ __forceinline__ bool clusterVisible(const FrameParams& fp, const RDyn& rd, const float4 cs, const float4 ca)
{
	const V3 c = affine(rd.mv, cs.x, cs.y, cs.z);
	const float R = cs.w * rd.radiusScale;
	const float slack = 1e-4f * (fabsf(c.x) + fabsf(c.y) + fabsf(c.z) + R) + 1e-6f;
	bool cull = c.z - R > fp.znear + slack;
#pragma unroll
	for (int k = 0; k < 4; k++)
		cull = cull || (fp.cullPlanes[k][0] * c.x + fp.cullPlanes[k][1] * c.y + fp.cullPlanes[k][2] * c.z + fp.cullPlanes[k][3] < -(R + slack));
	// The eye is the origin of view space; p ranges over the bounding sphere, n over the normal cone:
	// all back-facing <=> angle(eye->centre, axis) + cone half-angle (+ margin) + angular radius < 90 degrees
	if ((rd.cullFlags & 1) && ca.w < 1.5f)
	{
		const float d = sqrtf(c.x * c.x + c.y * c.y + c.z * c.z);
		if (d > R + slack)
		{
			const float inv = 1.0f / (rd.radiusScale * d); // the axis goes through the similarity's linear part
			const float tdot = (c.x * (rd.mv[0] * ca.x + rd.mv[1] * ca.y + rd.mv[2] * ca.z) + c.y * (rd.mv[4] * ca.x + rd.mv[5] * ca.y + rd.mv[6] * ca.z) +
			                    c.z * (rd.mv[8] * ca.x + rd.mv[9] * ca.y + rd.mv[10] * ca.z)) * inv;
			const float sinB = R / d, cosB = sqrtf(fmaxf(1.0f - sinB * sinB, 0.0f));
			const float cosA = sqrtf(fmaxf(1.0f - ca.w * ca.w, 0.0f));
			cull = cull || (tdot > ca.w * cosB + cosA * sinB + 1e-3f);
		}
	}
	return !cull;
}

// ------------------------------------------------------------------------------------------
// Triangle setup shared by the direct and the clipped path (Renderer.cpp:198-224).
// a,b,c = projected corners (pixel x, pixel y, view z, depth term). Returns false if rejected.
// ------------------------------------------------------------------------------------------
struct Setup
{
	float n1x, n1y, n2x, n2y;
	int x0, x1, y0, y1; // the reference's pixel loops: columns x0..x1, rows y0..y1 (the clamped bbox)
	uint32_t flags;
	uint32_t skip;      // MR_SKIP_*: outermost column / row of the loops that provably holds no covered pixel
};
#define MR_SKIP_L 1u
#define MR_SKIP_R 2u
#define MR_SKIP_T 4u
#define MR_SKIP_B 8u

// Tight scan. The reference's loops start at the pixel centre floor(min) + 0.5 and end at floor(max) + 0.5, i.e. they
// visit centres up to half a pixel OUTSIDE the triangle's bounding box on every side, where no pixel is covered - in
// exact arithmetic. In float arithmetic a centre farther than MR_TIGHT_DELTA outside the box is still rejected by the
// reference's own test (Renderer.cpp:245) whenever the triangle is not a sliver:
//   * a centre q at distance >= d outside the box has a true barycentric <= -d / (2 L), L = box extent + 1 >= |q - p_i|
//     (sum_i k_i (x_i - q_x) = 0 with every x_i - q_x in [d, L], sum k_i = 1, at most two k_i negative);
//   * the computed e1, e2, k0 differ from the true barycentrics by less than 2^-19 S^2 + 2^-14.9 S, S = L^2 / area:
//     area carries a relative error <= 2^-21 S (cancellation of two products of magnitude <= L^2), which scales every
//     barycentric (|k_i| <= sqrt(2) S inside the scanned box); row start and <= 64 chain additions add
//     <= 2^-16 S (each rounding is 2^-24 of a partial sum <= 2 |n| L, |n| L <= sqrt(2) S).
// With 2^-18 S^2 + 2^-12 S as the error bound (2x / 7x the above), the outermost column or row is skipped only when
// (2^-18 S^2 + 2^-12 S) * 2 L < MR_TIGHT_DELTA holds for the triangle; otherwise (slivers, huge triangles clipped by
// the frame) the full loops run. The chain still starts at the reference's first column (a skipped column costs its
// two additions, not its test). tests/test_gpu_invariance.py compares tight and full scans bit for bit.
#define MR_TIGHT_DELTA 0.0625f
#ifndef MR_RASTER_UNROLL
#define MR_RASTER_UNROLL 1 // the pixel loop of small triangles is short and branchy: unrolled copies only cost instruction cache
#endif
constexpr int kRasterUnroll = MR_RASTER_UNROLL;

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
	const float bx0 = minx, bx1 = maxx, by0 = miny, by1 = maxy; // the box before clamping
	minx = tclamp(minx, 0.0f, w - 1.0f);
	maxx = tclamp(maxx, 0.0f, w - 1.0f);
	miny = tclamp(miny, 0.0f, h - 1.0f);
	maxy = tclamp(maxy, 0.0f, h - 1.0f);
	// Pixel loops: x = floor(minx)+0.5, +1 ... while x <= maxx+0.5 (float sum), same in y.
	// All loop values are exact half-integers, so the last index is floor((max+0.5f) - 0.5f).
	const float fx0 = floorf(minx), fy0 = floorf(miny);
	s.x0 = (int)fx0;
	s.y0 = (int)fy0;
	const float xlim = maxx + 0.5f, ylim = maxy + 0.5f;
	int x1 = (int)floorf(xlim - 0.5f), y1 = (int)floorf(ylim - 0.5f);
	if ((float)x1 + 0.5f > xlim) x1--;
	if ((float)(x1 + 1) + 0.5f <= xlim) x1++;
	if ((float)y1 + 0.5f > ylim) y1--;
	if ((float)(y1 + 1) + 0.5f <= ylim) y1++;
	s.x1 = x1;
	s.y1 = y1;
	s.flags = 0u;
	s.skip = 0u;
	if (fp.tightScan)
	{
		const float L = fmaxf(bx1 - bx0, by1 - by0) + 1.0f;
		const float S = L * L * fabsf(i2a);
		if ((S * S * 3.814697265625e-6f + S * 2.44140625e-4f) * (2.0f * L) < MR_TIGHT_DELTA) // 2^-18, 2^-12; false for NaN
		{
			if (fx0 + 0.5f < bx0 - MR_TIGHT_DELTA) s.skip |= MR_SKIP_L;
			if ((float)x1 + 0.5f > bx1 + MR_TIGHT_DELTA) s.skip |= MR_SKIP_R;
			if (fy0 + 0.5f < by0 - MR_TIGHT_DELTA) s.skip |= MR_SKIP_T;
			if ((float)y1 + 0.5f > by1 + MR_TIGHT_DELTA) s.skip |= MR_SKIP_B;
		}
	}
	return true;
}

// The reference's inside test `!(e1 < 0 || e2 < 0 || k0 < 0)` (Renderer.cpp:245) as one three-way
// minimum: fminf drops NaN operands exactly where the reference's comparisons are false, and
// -0 < 0 is false in both forms.
__device__ __forceinline__ bool insideTest(float e1, float e2, float k0) { return !(fminf(fminf(e1, e2), k0) < 0.0f); }

// Depth of a covered pixel from its edge functions (Renderer.cpp:247-261); d = iz[] or zz[].
__device__ __forceinline__ float pixelDepth(int persp, float e1, float e2, float d0, float d1, float d2)
{
	const float k0 = 1.0f - e1 - e2;
	if (persp)
		return 1.0f / (k0 * d0 + e1 * d1 + e2 * d2); // :255
	return k0 * d0 + e1 * d1 + e2 * d2 + 0.0f * 1.0f; // :261
}

// 64-bit minimum without a return value (RED, fire and forget: the warp does not wait for L2)
__device__ __forceinline__ void redMin64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.min.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// Small triangle (bbox of at most MR_SMALL_AREA pixel centres): the reference's loops D/E
// (Renderer.cpp:238-269) run right here by the setup thread; the depth test of every covered pixel
// is one 64-bit atomicMin on the pixel's key in gkeys (fire-and-forget RED in L2). No binning, no
// second pass over the triangle. Returns whether any fragment was emitted (if not, the triangle can
// never own a pixel and needs no records).
#ifndef MR_SMALL_AREA
#define MR_SMALL_AREA 64
#endif
__device__ __forceinline__ bool rasterSmall(const FrameParams& fp, const float4 a, const float4 b, const float4 c, const Setup& s, int id)
{
	const int skipL = (int)(s.skip & MR_SKIP_L), skipR = (int)((s.skip >> 1) & 1u), skipT = (int)((s.skip >> 2) & 1u), skipB = (int)((s.skip >> 3) & 1u);
	const int ya = max(s.y0 + skipT, fp.rowBegin), yb = min(s.y1 - skipB, fp.rowEnd - 1);
	const int W = s.x1 - s.x0 + 1 - skipL - skipR; // columns tested per row
	const float ptx = (float)s.x0 + 0.5f;
	// row start (Renderer.cpp:241-242): e = n.x * (x0 - p.x) + n.y * (y - p.y); the first products do not depend on the row
	const float t1 = s.n1x * (ptx - c.x), t2 = s.n2x * (ptx - a.x);
	const unsigned long long idp1 = (unsigned long long)(uint32_t)(id + 1);
	bool any = false;
	if (W > 0)
	{
		float fy = (float)ya + 0.5f;
		unsigned long long* row = fp.gkeys + (size_t)ya * fp.w + (s.x0 + skipL);
		for (int y = ya; y <= yb; y++, fy += 1.0f, row += fp.w)
		{
			float e1 = t1 + s.n1y * (fy - c.y);
			float e2 = t2 + s.n2y * (fy - a.y);
			if (skipL) // the chain starts at the reference's first column whether or not that column is tested
			{
				e1 += s.n1x;
				e2 += s.n2x;
			}
#pragma unroll kRasterUnroll
			for (int i = 0; i < W; i++, e1 += s.n1x, e2 += s.n2x)
			{
				const float k0 = 1.0f - e1 - e2;
				if (insideTest(e1, e2, k0)) // Renderer.cpp:245
				{
					const float z = pixelDepth(fp.persp, e1, e2, a.w, b.w, c.w);
					if (z == z) // a NaN depth never passes `z < pixdepth`
					{
						redMin64(row + i, ((unsigned long long)zkey(z) << 32) | idp1);
						any = true;
					}
				}
			}
		}
	}
	return any;
}

// Where record `id` = 2 * triangle instance + clipper output lives. A record is 10 float4 fields (0-3
// raster record, 4-9 shading record) moved as 5 *pairs* of 32 bytes: Blackwell's 256-bit global loads and
// stores (LDG.256 / STG.256) carry one pair per instruction, and a pair is exactly one DRAM / L2 sector.
// Pair j (fields 2j, 2j+1) is at p[j * stride].
//  * sub-triangle 0 (every unclipped triangle): *plane* layout. The 32 triangles of a setup warp form
//    a block of 5 planes of 32 pairs, so each pair is written by one fully coalesced 1 KB store per warp,
//    and neighbouring winners read neighbouring sectors of the same lines.
//  * sub-triangle 1 (second clipper output, rare): plain 160-byte records in recs1.
struct __align__(32) F8
{
	float4 a, b;
};

// L2 eviction priority of the 256-bit accesses (PTX level2::eviction_priority, SASS .ENL2 / .ELL2 / .EFL2):
// records are written once by k_geom and read back by k_raster within the frame (keep: evict_last),
// finished tiles are never read again on the device (evict_first).
enum { L2_NORMAL = 0, L2_LAST = 1, L2_FIRST = 2 };
#ifndef MR_REC_STORE_HINT
#define MR_REC_STORE_HINT L2_LAST
#endif
#ifndef MR_REC_LOAD_HINT
#define MR_REC_LOAD_HINT L2_FIRST
#endif
#ifndef MR_TILE_STORE_HINT
#define MR_TILE_STORE_HINT L2_FIRST
#endif

template <int HINT = MR_REC_LOAD_HINT>
__device__ __forceinline__ F8 ldPair(const F8* p) // read-only path (records are written by an earlier kernel)
{
	F8 r;
	if (HINT == L2_FIRST)
		asm volatile("ld.global.nc.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
		             : "l"(p));
	else if (HINT == L2_LAST)
		asm volatile("ld.global.nc.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
		             : "l"(p));
	else
		asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
		             : "l"(p));
	return r;
}

// WIDE = false: two 16-byte stores. (ptxas 12.9 assembles st.global.v8 inside the out-of-line near-plane
// path as a plain 32-bit STG — found through the clipping parity test — so that rare path keeps float4 stores;
// minirender_b200/build.py counts the wide instructions in the SASS after every compile.)
template <bool WIDE, int HINT = L2_NORMAL>
__device__ __forceinline__ void stPair(F8* p, const float4 a, const float4 b)
{
	if (WIDE && HINT == L2_LAST)
		asm volatile("st.global.L2::evict_last.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
		             "f"(b.y), "f"(b.z), "f"(b.w)
		             : "memory");
	else if (WIDE && HINT == L2_FIRST)
		asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
		             "f"(b.y), "f"(b.z), "f"(b.w)
		             : "memory");
	else if (WIDE)
		asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
		             "f"(b.z), "f"(b.w)
		             : "memory");
	else
	{
		reinterpret_cast<float4*>(p)[0] = a;
		reinterpret_cast<float4*>(p)[1] = b;
	}
}

struct RecRef
{
	F8* p;
	int stride;
};

__device__ __forceinline__ RecRef recRef(const FrameParams& fp, int id)
{
	const int t = id >> 1;
	RecRef r;
	if (id & 1)
	{
		r.p = reinterpret_cast<F8*>(fp.recs1) + (size_t)t * (MR_REC_FIELDS / 2);
		r.stride = 1;
	}
	else
	{
		r.p = reinterpret_cast<F8*>(fp.recs) + (size_t)(t >> 5) * (MR_REC_FIELDS / 2 * 32) + (t & 31);
		r.stride = 32;
	}
	return r;
}

// One corner in view space: what paintMesh's loops A/B/C hand to paintTriangle (Renderer.cpp:344-380).
struct Corner
{
	float px, py, pz, nx, ny, nz, u, v;
};

// A record = 10 float4 fields in 5 pairs. With P = (view position, nx) and Q = (ny, nz, u, v) of a corner - the two
// planes k_geom keeps its transformed corners in:
//   pair 0  (a.x a.y c.x c.y)          (n1x n1y n2x n2y)              edge origins, scaled edge normals
//   pair 1  (a.w b.w c.w material)     (x0 | x1 << 16, nx0 nx1 nx2)   depth terms, material, column span
//   pair 2  (Q0.xy Q1.xy)              (Q2.xy, submission, pos2.z)    the other normal components
//   pair 3  (pos0.xyz pos1.x)          (pos1.yz pos2.xy)              corner positions: ONLY when a pixel's position cannot
//                                                                     be unprojected from its own ray (fp.unproject == 0)
//   pair 4  (Q0.zw Q1.zw)              (Q2.zw, y0 | y1 << 16, flags)  ONLY what textured shading, the binned raster and
//                                                                     the checkpoints need
// Pair 4 is written (`pair4`) when the frame textures, runs k_chain, or the triangle is binned / clipped, and read only
// by those consumers: under the standard perspective an untextured pixel is resolved with three 32-byte loads and a
// small untextured triangle stored with three.
__device__ __forceinline__ void storeRecClipped(const FrameParams& fp, const RecRef d, const float4 a, const float4 b, const float4 c, const Setup& s, int material,
                                                int submission, const Corner& o0, const Corner& o1, const Corner& o2)
{
	stPair<false, L2_NORMAL>(d.p, make_float4(a.x, a.y, c.x, c.y), make_float4(s.n1x, s.n1y, s.n2x, s.n2y));
	stPair<false, L2_NORMAL>(d.p + d.stride, make_float4(a.w, b.w, c.w, __uint_as_float((uint32_t)material)),
	                         make_float4(__uint_as_float((uint32_t)s.x0 | ((uint32_t)s.x1 << 16)), o0.nx, o1.nx, o2.nx));
	stPair<false, L2_NORMAL>(d.p + 2 * d.stride, make_float4(o0.ny, o0.nz, o1.ny, o1.nz), make_float4(o2.ny, o2.nz, __uint_as_float((uint32_t)submission), o2.pz));
	if (!fp.unproject)
		stPair<false, L2_NORMAL>(d.p + 3 * d.stride, make_float4(o0.px, o0.py, o0.pz, o1.px), make_float4(o1.py, o1.pz, o2.px, o2.py));
	stPair<false, L2_NORMAL>(d.p + 4 * d.stride, make_float4(o0.u, o0.v, o1.u, o1.v),
	                         make_float4(o2.u, o2.v, __uint_as_float((uint32_t)s.y0 | ((uint32_t)s.y1 << 16)), __uint_as_float(s.flags)));
}

// The same for a triangle whose corners lie in k_geom's shared-memory planes sP / sQ: every piece is fetched right before
// the pair that needs it (the stores are ordered asm statements, so the fetches stay where they are written).
__device__ __forceinline__ void storeRecCorners(const FrameParams& fp, const RecRef d, const float4 a, const float4 b, const float4 c, const Setup& s, int material,
                                                int submission, const float4* __restrict__ sP, const float4* __restrict__ sQ, int i0, int i1, int i2, bool pair4)
{
	stPair<true, MR_REC_STORE_HINT>(d.p, make_float4(a.x, a.y, c.x, c.y), make_float4(s.n1x, s.n1y, s.n2x, s.n2y));
	stPair<true, MR_REC_STORE_HINT>(d.p + d.stride, make_float4(a.w, b.w, c.w, __uint_as_float((uint32_t)material)),
	                                make_float4(__uint_as_float((uint32_t)s.x0 | ((uint32_t)s.x1 << 16)), sP[i0].w, sP[i1].w, sP[i2].w));
	{
		const float2 n0 = *reinterpret_cast<const float2*>(&sQ[i0]), n1 = *reinterpret_cast<const float2*>(&sQ[i1]), n2 = *reinterpret_cast<const float2*>(&sQ[i2]);
		stPair<true, MR_REC_STORE_HINT>(d.p + 2 * d.stride, make_float4(n0.x, n0.y, n1.x, n1.y),
		                                make_float4(n2.x, n2.y, __uint_as_float((uint32_t)submission), sP[i2].z));
	}
	if (!fp.unproject)
	{
		const float4 p0 = sP[i0], p1 = sP[i1], p2 = sP[i2];
		stPair<true, MR_REC_STORE_HINT>(d.p + 3 * d.stride, make_float4(p0.x, p0.y, p0.z, p1.x), make_float4(p1.y, p1.z, p2.x, p2.y));
	}
	if (pair4)
	{
		const float2 t0 = reinterpret_cast<const float2*>(&sQ[i0])[1], t1 = reinterpret_cast<const float2*>(&sQ[i1])[1], t2 = reinterpret_cast<const float2*>(&sQ[i2])[1];
		stPair<true, MR_REC_STORE_HINT>(d.p + 4 * d.stride, make_float4(t0.x, t0.y, t1.x, t1.y),
		                                make_float4(t2.x, t2.y, __uint_as_float((uint32_t)s.y0 | ((uint32_t)s.y1 << 16)), __uint_as_float(s.flags)));
	}
}

// Edge-chain checkpoints for a wide binned triangle (record `id`, pixel loops x0..x1, y0..y1): reserves rows x (tile
// boundaries inside the bbox) pool entries and one k_chain work item per block of 32 rows. Returns what goes into
// bits 1..31 of the record's flags word (1 + first pool entry), or 0 - not wide enough, no k_chain this frame, or
// the buffers are full - in which case the tile kernel walks the chain from the triangle's first column as before.
// What the frame would have needed is counted either way: the host sizes the next frames' buffers from it.
__device__ __noinline__ uint32_t chkReserve(const FrameParams& fp, int id, int x0, int x1, int y0, int y1)
{
	const int ntb = (x1 >> MR_TILE_SHIFT) - (x0 >> MR_TILE_SHIFT);
	if (ntb < fp.chkMinTiles || y1 < y0)
		return 0u;
	const int rows = y1 - y0 + 1, nblk = (rows + 31) >> 5;
	const unsigned n = (unsigned)rows * (unsigned)ntb;
	atomicAdd(&fp.ctr->chkDemand, (unsigned long long)n);
	atomicAdd(&fp.ctr->chkItemDemand, (unsigned)nblk);
	if (!fp.chkEnable)
		return 0u;
	const unsigned base = atomicAdd(&fp.ctr->chkUsed, n);
	const bool fits = (unsigned long long)base + n <= (unsigned long long)fp.chkCap && base + n < 0x7fffffffu;
	const unsigned it = atomicAdd(&fp.ctr->chkItems, (unsigned)nblk);
	const bool ok = fits && it + (unsigned)nblk <= (unsigned)fp.chkItemCap;
	for (int b = 0; b < nblk; b++)
		if (it + (unsigned)b < (unsigned)fp.chkItemCap)
			fp.chkItems[it + b] = make_int4(ok ? id : -1, (int)base, b, 0); // (a reserved item is always written: k_chain reads them all)
	return ok ? (base + 1u) << 1 : 0u;
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

// Appends record `id` to tile `tile`'s bin at position `slot` (from the tile counter); entries
// beyond the bin capacity go to the global overflow list that the tile kernel scans.
__device__ __forceinline__ void binStore(const FrameParams& fp, int tile, int slot, int id)
{
	if (slot < fp.binCap)
		fp.bins[(size_t)tile * fp.binCap + slot] = id;
	else
	{
		const unsigned long long o = atomicAdd(&fp.ctr->ovfTotal, 1ull);
		if (o < (unsigned long long)fp.ovfCap)
			fp.ovfPairs[o] = make_int2(tile, id);
		else
			fp.ctr->overflow = 1u;
	}
}

// Bins one record into every tile of its bbox, the whole warp working on it (32 tiles per step).
// All arguments are warp-uniform. Used for triangles spanning more than MR_SEG_PER_LANE tiles and
// for clipper output: a lone lane walking thousands of tiles is the critical path of scenes with
// huge triangles.
__device__ __forceinline__ void binCooperative(const FrameParams& fp, int lane, int id, int x0, int x1, int y0, int y1)
{
	const int tx0 = x0 >> MR_TILE_SHIFT, tx1 = x1 >> MR_TILE_SHIFT;
	const int ty0 = max(y0 >> MR_TILE_SHIFT, fp.tileRow0), ty1 = min(y1 >> MR_TILE_SHIFT, fp.tileRow0 + fp.tileRows - 1);
	const int nx = tx1 - tx0 + 1, n = nx * (ty1 - ty0 + 1);
	for (int k = lane; k < n; k += 32)
	{
		const int row = k / nx;
		const int tile = (ty0 + row) * fp.tilesX + tx0 + k - row * nx;
		binStore(fp, tile, atomicAdd(&fp.tileCount[tile].x, 1), id);
	}
}

__device__ __forceinline__ Corner cornerOf(const float4 B, const float4 C)
{
	Corner c;
	c.px = B.x; c.py = B.y; c.pz = B.z; c.nx = B.w; // k_geom's planes: (position, nx) (ny, nz, u, v)
	c.ny = C.x; c.nz = C.y; c.u = C.z; c.v = C.w;
	return c;
}

// Near-plane path of k_geom (rare): clips the triangle's view-space corners, sets up and stores the
// records of the one or two output triangles (the caller's warp bins them). Returns a bit per stored
// sub-triangle (bits 0, 1; bits 2, 3: k_chain bins that one). Out of line so that its stack never touches the fast path.
__device__ __noinline__ int setupClipped(const FrameParams& fp, int t, int material, int submission, const float4* sB, const float4* sC, int i0, int i1, int i2,
                                         uint2* spans)
{
	const Corner v0 = cornerOf(sB[i0], sC[i0]), v1 = cornerOf(sB[i1], sC[i1]), v2 = cornerOf(sB[i2], sC[i2]);
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
		s.flags = MR_REC_CLIPPED | chkReserve(fp, id, s.x0, s.x1, s.y0, s.y1);
		const RecRef ref = recRef(fp, id);
		storeRecClipped(fp, ref, a, b, c, s, material, submission + sub, o0, o1, o2);
		spans[sub] = make_uint2((uint32_t)s.x0 | ((uint32_t)s.x1 << 16), (uint32_t)s.y0 | ((uint32_t)s.y1 << 16));
		nrec |= (1 << sub) | ((s.flags >> 1) ? (4 << sub) : 0); // bits 2, 3: the sub-triangle has checkpoints and is binned by k_chain

	}
	return nrec;
}

// ------------------------------------------------------------------------------------------
// Kernel 1: geometry. Persistent CTAs of MR_GEOM_WARPS warps; in the main loop a warp works on its own: no
// block-wide barrier.
//
// Phase 0, cull: one thread per cluster of the frame (a cluster = MR_CLUSTER = 32 consecutive triangles of a
//   renderable; slices of 32 clusters are dealt to the warps of the grid round-robin, so every SM gets its share)
//   evaluates clusterVisible(); survivors are appended to the frame's work list as self-contained 128-byte
//   GeomEntry records (renderable matrices included), one global atomic per CTA. The CTAs then meet at a
//   grid-wide counter: the list is complete before anyone takes from it.
// Phase 1: a warp takes one entry at a time and processes that cluster entirely out of shared memory:
//     load      its meshlet blob (the cluster's distinct corners + 10-bit local indices, ~1.2 KB for a grid
//               mesh) arrives by one cp.async.bulk, completion on the warp's mbarrier; the copy for cluster
//               k+1 is issued as soon as cluster k's corners are transformed, so it overlaps k's triangles;
//     vertex    lane v transforms corner v (v + 32, ...): modelview (loop A, Renderer.cpp:344-345), htransform +
//               pixel coordinates + depth term (:13-20, :186-196, :223-224), normal matrix (loop B, :347-348);
//               results go to the warp's shared memory as three float4 planes;
//     triangle  lane j sets up triangle j from shared memory (near test, clip path, reject, cull, edge
//               normals, spans: Renderer.cpp:163-224). bbox of at most MR_SMALL_AREA pixel centres: the
//               reference's loops D/E run right here, the depth test being one 64-bit RED.MIN per covered
//               pixel on gkeys; larger: binned to the 16x16 tiles of its bbox for k_raster. A triangle that
//               can own a pixel writes its record (index 2t: the index is the submission id) as five
//               coalesced 32-byte pairs.
//   Work distribution: the list is cut into chunks of MR_GEOM_CHUNK entries. The warps of a CTA draw tickets from
//   a shared-memory counter; ticket t is entry t % CHUNK of the CTA's chunk number t / CHUNK. A CTA's first chunk
//   is fixed by its position; the warp that draws the first ticket of a later chunk fetches its number with ONE
//   global atomic (a single-address atomic per cluster would serialise in L2, ~5 cycles each: measured as the
//   whole kernel's bound) after the previous chunk's number has arrived, so a CTA's chunk numbers grow with its
//   tickets and the first ticket behind the end of the list tells a warp that it is done. A CTA never holds more
//   than one chunk (one cluster per warp) ahead of what it works on: balance across SMs stays dynamic to the end.
//   The entry of cluster k+1 is read while cluster k is processed and its bulk copy issued right after k's vertex
//   phase: no dependent global load sits on a warp's critical path.
// Everything a frame needs zeroed between frames (tile counters, list counters, the next frame's statistics)
// is reset by k_raster.
// The order in which fragments or bin entries arrive does not matter: depth ties are resolved on
// the record index (submission id) carried in the low word of every depth key.
// ------------------------------------------------------------------------------------------
#ifndef MR_GEOM_WARPS
#define MR_GEOM_WARPS 8 // x 2 CTAs per SM: as fast as twice as many (measured), without register spills, and a cluster takes half as long
#endif
#define MR_GEOM_THREADS (MR_GEOM_WARPS * 32)
#ifndef MR_GEOM_MINB
#define MR_GEOM_MINB 2
#endif
static_assert(MR_CLUSTER == 32, "a cluster is what one warp sets up at a time");
#define MR_GEOM_CHUNK 8
#define MR_GEOM_RING 4

struct GeomShared // fixed part of k_geom's dynamic shared memory
{
	GeomEntry entry[MR_GEOM_WARPS][2];      // the cluster a warp works on (k & 1) and the next one
	unsigned long long full[MR_GEOM_WARPS]; // mbarrier: the warp's meshlet (or the news that there is none) has arrived
	unsigned long long ring[MR_GEOM_RING];  // chunk slot number << 32 | chunk number
	unsigned long long stat;
	unsigned ticket;
	int ownCount;            // warps of this CTA that start on a cluster the CTA found in its own first cull round
	__align__(16) unsigned dirty[MR_GEOM_WARPS][4]; // per warp: tiles it may have drawn into (Counters::dirty encoding)
	int cullCount, cullBase; // survivors of this CTA in the current cull round, and where they go in the list
};
#define MR_GEOM_FIXED_BYTES ((int)((sizeof(GeomShared) + 127) / 128 * 128))
// per warp: the meshlet as loaded, the transformed corners (3 float4 planes), a copy of the local indices
#define MR_GEOM_WARP_BYTES(nvCap) (MR_MESHLET_BYTES(nvCap) + (nvCap) * 48 + MR_CLUSTER * 4)

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(unsigned long long* b, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(b)) : "memory"); }
__device__ __forceinline__ void mbarWait(unsigned long long* b, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"MR_WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra MR_DONE_%=;\n"
		"bra MR_WAIT_%=;\n"
		"MR_DONE_%=:\n"
		"}" ::"r"(smemAddr(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, non-tensor form), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)), "l"(src), "r"(bytes),
	             "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ int ldAcquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// The next work-list index for the calling lane (lane 0 of a warp), or 0x7fffffff when the list is exhausted; see k_geom.
// The warps of a CTA draw tickets from a shared-memory counter. The first tickets of a CTA are fixed by its position
// (ticket t of CTA b = entry b + t * grid: interleaved over the list, no global access at all); only the last
// MR_GEOM_DYNAMIC_TAIL entries per CTA of the list are handed out by a global counter, in chunks of MR_GEOM_CHUNK
// (a single-address atomic per cluster would serialise in L2), where the balance across SMs is decided.
#ifndef MR_GEOM_DYNAMIC_TAIL
#define MR_GEOM_DYNAMIC_TAIL 8
#endif
__device__ __forceinline__ int geomStatic(int nVis, int grid) { return max(nVis / grid - MR_GEOM_DYNAMIC_TAIL, 0); } // positional entries per CTA
__device__ __forceinline__ int geomPop(GeomShared& gs, int* sync, int nVis, int grid, unsigned nStatic /* geomStatic(nVis, grid): a division, once per warp */)
{
	const unsigned t = atomicAdd(&gs.ticket, 1u);
	if (t < nStatic)
		return (int)blockIdx.x + (int)t * grid; // (< nStatic * grid <= nVis)
	const unsigned td = t - nStatic, slot = td / MR_GEOM_CHUNK, within = td % MR_GEOM_CHUNK;
	volatile unsigned long long* ring = gs.ring;
	if (within == 0u)
	{
		// this warp fetches the chunk for everybody, once the previous chunk's number is known (keeps them ordered)
		if (slot > 0u)
			while ((unsigned)(ring[(slot - 1u) % MR_GEOM_RING] >> 32) != slot - 1u)
				;
		const int c = atomicAdd(&sync[0], 1);
		ring[slot % MR_GEOM_RING] = ((unsigned long long)slot << 32) | (unsigned)c;
	}
	unsigned long long v = ring[slot % MR_GEOM_RING];
	while ((unsigned)(v >> 32) != slot)
		v = ring[slot % MR_GEOM_RING];
	const long long idx = (long long)nStatic * grid + (long long)(unsigned)v * MR_GEOM_CHUNK + within;
	return idx < (long long)nVis ? (int)idx : 0x7fffffff;
}

// Grows the warp's rectangle of tiles that may have been drawn into (lane 0 calls this; `dirty` is the warp's own
// 16-byte slot in shared memory: a plain read-modify-write, no atomics).
__device__ __forceinline__ void markDirty(const FrameParams& fp, unsigned* dirty, int tx0, int tx1, int ty0, int ty1)
{
	uint4 cur = *reinterpret_cast<uint4*>(dirty);
	cur.x = max(cur.x, (unsigned)(fp.tilesX - max(tx0, 0)));
	cur.y = max(cur.y, (unsigned)(min(tx1, fp.tilesX - 1) + 1));
	cur.z = max(cur.z, (unsigned)(fp.tilesY - max(ty0, 0)));
	cur.w = max(cur.w, (unsigned)(min(ty1, fp.tilesY - 1) + 1));
	*reinterpret_cast<uint4*>(dirty) = cur;
}

// Binning of one warp's larger triangles (`binned` lanes: up to MR_SEG_PER_LANE tiles each, warp-aggregated; more:
// the whole warp, one triangle at a time) and of the clipper's output triangles. Called by all lanes of the warp.
__device__ __noinline__ void geomBin(const FrameParams& fp, int lane, int t, bool binned, int nrecSlow, int sx0, int sx1, int sy0, int sy1, uint2 clip0, uint2 clip1,
                                     unsigned* dirty)
{
	{
		// the tiles these triangles may draw into, for the frame's dirty rectangle
		int x0 = binned ? sx0 : 0x7fff, x1 = binned ? sx1 : -1, y0 = binned ? sy0 : 0x7fff, y1 = binned ? sy1 : -1;
		if (nrecSlow & 1)
		{
			x0 = min(x0, (int)(clip0.x & 0xffffu)); x1 = max(x1, (int)(clip0.x >> 16));
			y0 = min(y0, (int)(clip0.y & 0xffffu)); y1 = max(y1, (int)(clip0.y >> 16));
		}
		if (nrecSlow & 2)
		{
			x0 = min(x0, (int)(clip1.x & 0xffffu)); x1 = max(x1, (int)(clip1.x >> 16));
			y0 = min(y0, (int)(clip1.y & 0xffffu)); y1 = max(y1, (int)(clip1.y >> 16));
		}
		x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
		y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
		if (lane == 0 && x1 >= x0)
			markDirty(fp, dirty, x0 >> MR_TILE_SHIFT, x1 >> MR_TILE_SHIFT, max(y0, fp.rowBegin) >> MR_TILE_SHIFT, min(y1, fp.rowEnd - 1) >> MR_TILE_SHIFT);
	}
	// ---- wide triangles get checkpoints of their edge chains (k_chain): the first pool entry goes into the flags word of
	// the record this lane has just stored (same thread, same address: ordered behind that store) ----
	if (binned && (sx1 >> MR_TILE_SHIFT) - (sx0 >> MR_TILE_SHIFT) >= fp.chkMinTiles)
	{
		const uint32_t f = chkReserve(fp, 2 * t, sx0, sx1, sy0, sy1);
		if (f != 0u)
		{
			const RecRef ref = recRef(fp, 2 * t);
			reinterpret_cast<uint32_t*>(ref.p + 4 * ref.stride)[7] = f; // pair 4 = (. . . . | . . rows FLAGS)
			binned = false; // k_chain bins it as well: one warp per 32 rows instead of this warp for the whole triangle
		}
	}
	// ---- binning of the larger triangles: up to MR_SEG_PER_LANE tiles each, warp-aggregated ----
	if (__any_sync(0xffffffffu, binned))
	{
		const int tx0 = sx0 >> MR_TILE_SHIFT, tx1 = sx1 >> MR_TILE_SHIFT;
		const int ty0 = max(sy0 >> MR_TILE_SHIFT, fp.tileRow0), ty1 = min(sy1 >> MR_TILE_SHIFT, fp.tileRow0 + fp.tileRows - 1);
		const int nx = tx1 - tx0 + 1;
		const int ntiles = binned ? nx * (ty1 - ty0 + 1) : 0;
		const bool big = ntiles > MR_SEG_PER_LANE;
		const int mine = big ? 0 : ntiles;
		const int id = 2 * t;
		const int rounds = __reduce_max_sync(0xffffffffu, mine);
		// issue the counter atomics of all rounds first, use their results afterwards
		int tileOf[MR_SEG_PER_LANE], baseOf[MR_SEG_PER_LANE], rankOf[MR_SEG_PER_LANE], leadOf[MR_SEG_PER_LANE];
#pragma unroll
		for (int k = 0; k < MR_SEG_PER_LANE; k++)
		{
			tileOf[k] = -1; baseOf[k] = 0; rankOf[k] = 0; leadOf[k] = 0;
			if (k < rounds)
			{
				const bool on = k < mine;
				const int krow = (k >= nx) + (k >= 2 * nx) + (k >= 3 * nx); // k / nx for k < 4
				const int tile = on ? (ty0 + krow) * fp.tilesX + tx0 + k - krow * nx : -1 - lane;
				const unsigned peers = __match_any_sync(0xffffffffu, tile);
				leadOf[k] = __ffs(peers) - 1;
				rankOf[k] = __popc(peers & ((1u << lane) - 1u));
				if (on)
				{
					tileOf[k] = tile;
					if (lane == leadOf[k])
						baseOf[k] = atomicAdd(&fp.tileCount[tile].x, __popc(peers));
				}
			}
		}
#pragma unroll
		for (int k = 0; k < MR_SEG_PER_LANE; k++)
			if (k < rounds)
			{
				const int slot = __shfl_sync(0xffffffffu, baseOf[k], leadOf[k]) + rankOf[k];
				if (tileOf[k] >= 0)
					binStore(fp, tileOf[k], slot, id);
			}
		// Triangles spanning more tiles: the warp bins them together, one at a time.
		unsigned bigLanes = __ballot_sync(0xffffffffu, big);
		while (bigLanes != 0u)
		{
			const int src = __ffs(bigLanes) - 1;
			bigLanes &= bigLanes - 1u;
			binCooperative(fp, lane, 2 * (t - lane + src), __shfl_sync(0xffffffffu, sx0, src), __shfl_sync(0xffffffffu, sx1, src),
			               __shfl_sync(0xffffffffu, sy0, src), __shfl_sync(0xffffffffu, sy1, src));
		}
	}
	// clipper output: binned by the whole warp as well
	unsigned clipLanes = __ballot_sync(0xffffffffu, nrecSlow != 0);
	while (clipLanes != 0u)
	{
		const int src = __ffs(clipLanes) - 1;
		clipLanes &= clipLanes - 1u;
		const int subs = __shfl_sync(0xffffffffu, nrecSlow, src);
		for (int sub = 0; sub < 2; sub++)
			if ((subs & (1 << sub)) && !(subs & (4 << sub)))
			{
				const int id = 2 * (t - lane + src) + sub;
				const uint32_t xs = __shfl_sync(0xffffffffu, (sub ? clip1.x : clip0.x), src), ys = __shfl_sync(0xffffffffu, (sub ? clip1.y : clip0.y), src);
				binCooperative(fp, lane, id, xs & 0xffffu, xs >> 16, ys & 0xffffu, ys >> 16);
			}
	}

}

// Triangle phase of one cluster for one lane (triangle `lane` of the cluster); see k_geom.
// acc = this thread's statistics: records | clipped inputs << 20 | zero-coverage drops << 40.
__device__ __forceinline__ void geomTriangle(const FrameParams& fp, const GeomEntry& e, const float4* __restrict__ sA, const float4* __restrict__ sB,
                                             const float4* __restrict__ sC, const uint32_t* __restrict__ sIdx, int lane, unsigned long long& acc, unsigned* dirty)
{
	const int t = e.ci * MR_CLUSTER + lane; // triangle instance (padded numbering): 2t is its record index
	const int tri = e.triFirst + lane;      // triangle within the mesh
	const bool active = tri < e.nTri;     // false: padding behind the mesh's last triangle
	bool valid = false, binned = false;
	int nclip = 0, nrecSlow = 0, nzero = 0;
	uint2 clipSpans[2]; // bbox spans of the clipper's output triangles (the warp bins them below)
	clipSpans[0] = clipSpans[1] = make_uint2(0u, 0u);
	Setup s;
	s.x0 = s.x1 = s.y0 = s.y1 = 0;
	s.flags = 0u;
	if (active)
	{
		const uint32_t ix = sIdx[lane];
		const int i0 = ix & 1023u, i1 = (ix >> 10) & 1023u, i2 = ix >> 20;
		const float4 a = sA[i0], b = sA[i1], c = sA[i2];
		const float zn = fp.znear;
		if (a.z > zn || b.z > zn || c.z > zn) // Renderer.cpp:169-177
		{
			if (!(a.z > zn && b.z > zn && c.z > zn))
			{
				nclip = 1;
				nrecSlow = setupClipped(fp, t, e.material, 2 * (e.subBase + tri), sB, sC, i0, i1, i2, clipSpans);
			}
		}
		else if (setupTriangle(fp, a, b, c, s))
		{
			valid = min(s.y1 >> MR_TILE_SHIFT, fp.tileRow0 + fp.tileRows - 1) >= max(s.y0 >> MR_TILE_SHIFT, fp.tileRow0);
			if (valid)
			{
				if ((s.x1 - s.x0 + 1) * (s.y1 - s.y0 + 1) <= MR_SMALL_AREA)
				{
					valid = rasterSmall(fp, a, b, c, s, 2 * t);
					nzero = valid ? 0 : 1;
				}
				else
					binned = true;
			}
			if (valid)
			{
				// The triangle can own pixels: its record = raster half + its three view-space corners as they lie in shared memory
				const RecRef ref = recRef(fp, 2 * t);
				storeRecCorners(fp, ref, a, b, c, s, e.material, 2 * (e.subBase + tri), sB, sC, i0, i1, i2, binned || fp.texturing || fp.chkEnable || (fp.debug & 512));
			}
		}
	}
	__syncwarp();

	// ---- tiles that may have received fragments of this warp's small triangles: the tiles of the union
	// of their bboxes get their "touched" word set (plain idempotent stores, a handful per warp). The tile
	// kernel does not even read the depth keys of a tile nothing has touched. ----
	{
		const bool emitted = valid && !binned;
		const int ya = max(s.y0, fp.rowBegin), yb = min(s.y1, fp.rowEnd - 1);
		const int tx0 = __reduce_min_sync(0xffffffffu, emitted ? (s.x0 >> MR_TILE_SHIFT) : 0x7fff);
		const int tx1 = __reduce_max_sync(0xffffffffu, emitted ? (s.x1 >> MR_TILE_SHIFT) : -1);
		const int ty0 = __reduce_min_sync(0xffffffffu, emitted ? (ya >> MR_TILE_SHIFT) : 0x7fff);
		const int ty1 = __reduce_max_sync(0xffffffffu, emitted ? (yb >> MR_TILE_SHIFT) : -1);
		if (tx1 >= tx0)
		{
			if (lane == 0)
				markDirty(fp, dirty, tx0, tx1, ty0, ty1);
			// 8 x 4 tiles per pass; a warp of neighbouring small triangles spans a few tiles: one pass
			const int tx = tx0 + (lane & 7), ty = ty0 + (lane >> 3);
			if (tx1 - tx0 < 8 && ty1 - ty0 < 4)
			{
				if (tx <= tx1 && ty <= ty1)
					fp.tileCount[ty * fp.tilesX + tx].y = 1;
			}
			else
				for (int y = ty; y <= ty1; y += 4)
					for (int x = tx; x <= tx1; x += 8)
						fp.tileCount[y * fp.tilesX + x].y = 1;
		}
	}

	// ---- the larger triangles and the clipper's output go to the tile bins (rare on fine meshes: out of line, so that
	// the code a warp normally runs stays small) ----
	if (__any_sync(0xffffffffu, binned || nrecSlow != 0))
		geomBin(fp, lane, t, binned, nrecSlow, s.x0, s.x1, s.y0, s.y1, clipSpans[0], clipSpans[1], dirty);

	acc += (unsigned long long)((valid ? 1 : 0) + __popc(nrecSlow & 3)) | ((unsigned long long)nclip << 20) | ((unsigned long long)nzero << 40);
}

#ifdef MR_TIMELINE
// Experiment build only: per-warp time stamps (globaltimer, ns) of k_geom's phases, read back by mr_debug_timeline().
#define MR_TL_SLOTS 8
__device__ unsigned long long g_timeline[1024 * MR_GEOM_WARPS * MR_TL_SLOTS];
__device__ unsigned long long g_tlTri[1024 * MR_GEOM_WARPS]; // k_geom: per warp, time in the triangle phase
__device__ unsigned long long g_rtl[8192 * 4]; // k_raster: per tile CTA start, dependency resolved, end, SM
__device__ __forceinline__ unsigned long long globalTimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MR_TL(slot) do { if (lane == 0 && blockIdx.x < 1024) g_timeline[((size_t)blockIdx.x * MR_GEOM_WARPS + warp) * MR_TL_SLOTS + (slot)] = globalTimer(); } while (0)
#define MR_TL_ADD(slot, v) do { if (lane == 0 && blockIdx.x < 1024) g_timeline[((size_t)blockIdx.x * MR_GEOM_WARPS + warp) * MR_TL_SLOTS + (slot)] += (v); } while (0)
#else
#define MR_TL(slot) do { } while (0)
#define MR_TL_ADD(slot, v) do { } while (0)
#endif

template <int TM>
__global__ void __launch_bounds__(MR_GEOM_THREADS, MR_GEOM_MINB) k_geom(const __grid_constant__ FrameParams fp)
{
	extern __shared__ __align__(128) unsigned char geomSmem[];
	GeomShared& gs = *reinterpret_cast<GeomShared*>(geomSmem);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int nvCap = fp.geomVertCap;
	const int nCl = fp.nTriInst / MR_CLUSTER;
	int* const syncChunks = fp.geomSync;
	int* const syncTail = fp.geomSync + MR_SYNC_STRIDE;
	int* const syncArrived = fp.geomSync + 2 * MR_SYNC_STRIDE;

	pdlLaunchDependents(); // k_raster's CTAs may take the SMs this kernel's CTAs leave (they wait for the whole grid)
	MR_TL(0); // kernel entered
	if (tid == 0)
	{
		for (int i = 0; i < MR_GEOM_WARPS; i++)
			mbarInit(&gs.full[i], 1);
		gs.stat = 0ull;
		gs.ticket = 0u;
		gs.cullCount = 0;
		gs.ownCount = 0;
		for (int i = 0; i < MR_GEOM_WARPS * 4; i++)
			(&gs.dirty[0][0])[i] = 0u;
		for (int i = 0; i < MR_GEOM_RING; i++)
			gs.ring[i] = ~0ull;
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	// ---- phase 0: cull (Renderer.cpp:169-177, :202, :205-210 decided per cluster). Slice s = clusters 32 s .. 32 s + 31
	// goes to warp (s / grid) % WARPS of CTA s % grid, so a CTA looks at a few slices from distant parts of the frame
	// and nearly always finds visible clusters among them. The first MR_GEOM_WARPS survivors of its first round STAY
	// in the CTA - entry k straight into warp k's shared-memory slot, meshlet requested at once - as the warps' first
	// units of work; all others are appended to the frame's work list. The grid-wide rendezvous that completes the
	// list is then only needed when a warp takes its first entry from it, one cluster later: nobody waits for it. ----
	unsigned char* const warpMem = geomSmem + MR_GEOM_FIXED_BYTES + (size_t)warp * MR_GEOM_WARP_BYTES(nvCap);
	unsigned char* const raw = warpMem;
	unsigned long long* const full = &gs.full[warp];
	const int nSlices = (nCl + 31) >> 5;
	for (int round = 0; round * (int)gridDim.x * MR_GEOM_WARPS < nSlices; round++) // (block-uniform)
	{
		const int slice = (round * MR_GEOM_WARPS + warp) * (int)gridDim.x + (int)blockIdx.x;
		const int ci = slice * 32 + lane;
		bool vis = false;
		int r = 0;
		RStat rs;
		rs.triBase = rs.clusterBase = rs.nTri = rs.triBaseReal = 0;
		MeshletDir d;
		d.off16 = d.nv = 0u;
		if (slice < nSlices && ci < nCl)
		{
			r = (fp.nRenderables == 1) ? 0 : __ldg(&fp.triBlockCl[ci]);
			rs = frameRstat<TM>(fp)[r];
			const int meshCluster = rs.clusterBase + (ci * MR_CLUSTER - rs.triBase) / MR_CLUSTER;
			const float4* cl = fp.clusters + 2 * (size_t)meshCluster;
			const float4 cs = __ldg(cl), ca = __ldg(cl + 1);
			d = fp.meshletDir[meshCluster]; // requested together with the bounds
			vis = !fp.cullClusters || clusterVisible(fp, frameRdyn<TM>(fp)[r], cs, ca);
		}
		// where the survivors go: a shared-memory atomic per warp, one global atomic per CTA
		const unsigned m = __ballot_sync(0xffffffffu, vis);
		int at = 0;
		if (lane == 0 && m != 0u)
			at = atomicAdd(&gs.cullCount, __popc(m));
		at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
		__syncthreads();
		// (the position of the round's other survivors in the list is on its way while the kept entries are written)
		const int nSurv = gs.cullCount;
		const int keep = (round == 0 && !(fp.debug & 64)) ? min(nSurv, MR_GEOM_WARPS) : 0;
		int listBase = 0;
		if (tid == 0 && nSurv > keep)
			listBase = atomicAdd(syncTail, nSurv - keep);
		uint4 e0, e1, mv[3], nm[3];
		if (vis)
		{
			const RDyn& rd = frameRdyn<TM>(fp)[r];
			e0 = make_uint4((uint32_t)ci, d.nv & 0xffffu, d.off16, (uint32_t)(ci * MR_CLUSTER - rs.triBase));
			e1 = make_uint4((uint32_t)rs.nTri, (uint32_t)rs.triBaseReal, (uint32_t)rd.material, d.nv >> 16);
#pragma unroll
			for (int k = 0; k < 3; k++)
			{
				mv[k] = make_uint4(__float_as_uint(rd.mv[4 * k]), __float_as_uint(rd.mv[4 * k + 1]), __float_as_uint(rd.mv[4 * k + 2]), __float_as_uint(rd.mv[4 * k + 3]));
				nm[k] = make_uint4(__float_as_uint(rd.nm[4 * k]), __float_as_uint(rd.nm[4 * k + 1]), __float_as_uint(rd.nm[4 * k + 2]), __float_as_uint(rd.nm[4 * k + 3]));
			}
			if (at < keep)
			{
				uint4* o = reinterpret_cast<uint4*>(&gs.entry[at][0]);
				o[0] = e0; o[1] = e1;
#pragma unroll
				for (int k = 0; k < 3; k++)
				{
					o[2 + k] = mv[k];
					o[5 + k] = nm[k];
				}
			}
		}
		if (round == 0)
		{
			__syncthreads();
			if (warp < keep && lane == 0)
			{
				// this warp's first cluster: its meshlet is requested before the list is even written
				const GeomEntry& n0 = gs.entry[warp][0];
				mbarExpectTx(full, (uint32_t)MR_MESHLET_BYTES(n0.nv));
				bulkLoad(raw, fp.meshlets + (size_t)n0.off16 * 16, (uint32_t)MR_MESHLET_BYTES(n0.nv), full);
			}
		}
		if (tid == 0)
		{
			if (round == 0)
				gs.ownCount = keep;
			gs.cullBase = listBase;
			gs.cullCount = 0;
		}
		__syncthreads();
		if (vis && at >= keep)
		{
			uint4* o = reinterpret_cast<uint4*>(fp.visEntries + gs.cullBase + (at - keep));
			o[0] = e0; o[1] = e1;
#pragma unroll
			for (int k = 0; k < 3; k++)
			{
				o[2 + k] = mv[k];
				o[5 + k] = nm[k];
			}
		}
	}
	// this CTA's entries are published: arrive at the grid-wide counter (all CTAs are resident: the grid is sized from
	// the occupancy of this kernel and launched cooperatively). Who needs the list waits for the others there.
	// (the barrier orders every thread's entries before thread 0's fence, whose cumulativity publishes them with the
	// arrival: the other warps do not wait for the stores to reach L2)
	// (every thread fences its own entries: measured faster than one cumulative fence by thread 0 after the barrier,
	// which lets the warps start 1 us earlier but in lock-step - 38.7 vs 39.7 us on the sphere)
	__threadfence();
	MR_TL(1); // own clusters culled
	__syncthreads();
	if (tid == 0)
		atomicAdd(syncArrived, 1);
	const bool haveOwn = warp < gs.ownCount; // (warp-uniform) this warp starts on a cluster its CTA found itself

	// ---- phase 1: every warp on its own ----
	unsigned long long acc = 0ull;
	int nVis = -1; // (warp-uniform) length of the work list, known once every CTA has arrived
	int nStatic = 0; // (lane 0) entries per CTA dealt by position
	{
		float4* const sA = reinterpret_cast<float4*>(warpMem + MR_MESHLET_BYTES(nvCap));
		float4* const sB = sA + nvCap;
		float4* const sC = sB + nvCap;
		uint32_t* const sIdx = reinterpret_cast<uint32_t*>(sC + nvCap);
		const uint32_t* const visWords = reinterpret_cast<const uint32_t*>(fp.visEntries);
		const uint32_t none = (lane == 0) ? 0xffffffffu : 0u; // word `lane` of an entry that says "no more work"
		const int grid = (int)gridDim.x;
		uint32_t entryWord = none; // word `lane` of the next cluster's entry
		bool nextPending = haveOwn;  // the warp started on its own cluster: its next entry is fetched during that one
		if (!haveOwn)
		{
			int i0 = 0, i1 = 0, n = 0;
			if (lane == 0)
			{
				while (ldAcquire(syncArrived) < (int)gridDim.x)
					;
				n = ldAcquire(syncTail);
				nStatic = geomStatic(n, grid);
				i0 = geomPop(gs, syncChunks, n, grid, (unsigned)nStatic);
				i1 = (i0 < n) ? geomPop(gs, syncChunks, n, grid, (unsigned)nStatic) : i0;
			}
			nVis = __shfl_sync(0xffffffffu, n, 0);
			i0 = __shfl_sync(0xffffffffu, i0, 0);
			i1 = __shfl_sync(0xffffffffu, i1, 0);
			MR_TL(2); // work list complete
			const uint32_t w0 = (i0 < nVis) ? __ldcg(visWords + (size_t)i0 * 32 + lane) : none;
			entryWord = (i1 < nVis) ? __ldcg(visWords + (size_t)i1 * 32 + lane) : none;
			reinterpret_cast<uint32_t*>(&gs.entry[warp][0])[lane] = w0;
			__syncwarp();
			if (lane == 0)
			{
				const GeomEntry& n0 = gs.entry[warp][0];
				if (n0.ci >= 0)
				{
					mbarExpectTx(full, (uint32_t)MR_MESHLET_BYTES(n0.nv));
					bulkLoad(raw, fp.meshlets + (size_t)n0.off16 * 16, (uint32_t)MR_MESHLET_BYTES(n0.nv), full);
				}
			}
		}
#ifdef MR_TIMELINE
		else
			MR_TL(2);
		if (lane == 0 && blockIdx.x < 1024)
			g_timeline[((size_t)blockIdx.x * MR_GEOM_WARPS + warp) * MR_TL_SLOTS + 5] = g_timeline[((size_t)blockIdx.x * MR_GEOM_WARPS + warp) * MR_TL_SLOTS + 6] = g_timeline[((size_t)blockIdx.x * MR_GEOM_WARPS + warp) * MR_TL_SLOTS + 7] = g_tlTri[(size_t)blockIdx.x * MR_GEOM_WARPS + warp] = 0ull;
#endif
		for (int k = 0;; k++)
		{
			const GeomEntry& e = gs.entry[warp][k & 1];
			if (e.ci < 0)
				break;
#ifdef MR_TIMELINE
			const unsigned long long tw0 = globalTimer();
#endif
			mbarWait(full, (uint32_t)(k & 1));
#ifdef MR_TIMELINE
			MR_TL_ADD(5, globalTimer() - tw0); // time spent waiting for meshlets
			if (k == 0) MR_TL(3);              // first meshlet there
			MR_TL_ADD(6, 1ull);                // iterations
#endif
			const int nv = e.nv;
#ifdef MR_TIMELINE
			const unsigned long long tv0 = globalTimer();
#endif

			// ---- vertex phase: loops A and B of paintMesh for the cluster's corners, plus their projection ----
			{
				const float4* const raw0 = reinterpret_cast<const float4*>(raw);
				const float4* const raw1 = raw0 + nv;
				sIdx[lane] = reinterpret_cast<const uint32_t*>(raw1 + nv)[lane];
				if (fp.stdProj && !(e.flags & MR_MESHLET_NONFINITE))
					for (int v = lane; v < nv; v += 32)
					{
						const float4 q0 = raw0[v], q1 = raw1[v]; // (px py pz nx) (ny nz u v)
						const V3 view = affine(e.mv, q0.x, q0.y, q0.z);
						const V3 nrm = affine(e.nm, q0.w, q1.x, q1.y);
						sA[v] = projectStd(fp, view);
						sB[v] = make_float4(view.x, view.y, view.z, nrm.x); // planes as the records want them:
						sC[v] = make_float4(nrm.y, nrm.z, q1.z, q1.w);      // (position, nx) (ny, nz, u, v)
					}
				else
					for (int v = lane; v < nv; v += 32)
					{
						const float4 q0 = raw0[v], q1 = raw1[v];
						const V3 view = affine(e.mv, q0.x, q0.y, q0.z);
						const V3 nrm = affine(e.nm, q0.w, q1.x, q1.y);
						sA[v] = project(fp, view);
						sB[v] = make_float4(view.x, view.y, view.z, nrm.x); // planes as the records want them:
						sC[v] = make_float4(nrm.y, nrm.z, q1.z, q1.w);      // (position, nx) (ny, nz, u, v)
					}
			}
#ifdef MR_TIMELINE
			__syncwarp();
			MR_TL_ADD(7, globalTimer() - tv0); // time in the vertex phase
#endif
			// ---- the next cluster: its entry (read one iteration ago) into shared memory, its bulk copy issued
			// (the meshlet buffer is free again), another index drawn and that entry requested ----
			if (nextPending)
			{
				// first cluster of a warp that started on its own one: by now every CTA has long arrived
				int i1 = 0, n = 0;
				if (lane == 0)
				{
					while (ldAcquire(syncArrived) < (int)gridDim.x)
						;
					n = ldAcquire(syncTail);
					nStatic = geomStatic(n, grid);
					i1 = geomPop(gs, syncChunks, n, grid, (unsigned)nStatic);
				}
				nVis = __shfl_sync(0xffffffffu, n, 0);
				i1 = __shfl_sync(0xffffffffu, i1, 0);
				entryWord = (i1 < nVis) ? __ldcg(visWords + (size_t)i1 * 32 + lane) : none;
				nextPending = false;
			}
			reinterpret_cast<uint32_t*>(&gs.entry[warp][(k + 1) & 1])[lane] = entryWord;
			__syncwarp(); // corners complete, every lane is done with the meshlet buffer
			int i2 = 0x7fffffff;
			if (lane == 0)
			{
				const GeomEntry& n = gs.entry[warp][(k + 1) & 1];
				if (n.ci >= 0)
				{
					mbarExpectTx(full, (uint32_t)MR_MESHLET_BYTES(n.nv));
					bulkLoad(raw, fp.meshlets + (size_t)n.off16 * 16, (uint32_t)MR_MESHLET_BYTES(n.nv), full);
					i2 = geomPop(gs, syncChunks, nVis, grid, (unsigned)nStatic);
				}
			}
			i2 = __shfl_sync(0xffffffffu, i2, 0);
			entryWord = (i2 < nVis) ? __ldcg(visWords + (size_t)i2 * 32 + lane) : none;

			// ---- triangle phase ----
#ifdef MR_TIMELINE
			const unsigned long long tt0 = globalTimer();
#endif
			geomTriangle(fp, e, sA, sB, sC, sIdx, lane, acc, gs.dirty[warp]);
#ifdef MR_TIMELINE
			__syncwarp();
			if (lane == 0 && blockIdx.x < 1024) g_tlTri[(size_t)blockIdx.x * MR_GEOM_WARPS + warp] += globalTimer() - tt0;
#endif
			__syncwarp(); // the corners may be overwritten by the next cluster's
		}
	}

	MR_TL(4); // warp out of work
	{
		// k_raster starts every tile with a read of its counters; a cold L2 (they are reset, not rewritten, where nothing
		// was drawn) would make that a DRAM round trip per tile: the warps that are done here pull the array into L2
		const char* tc = reinterpret_cast<const char*>(fp.tileCount);
		const size_t bytes = sizeof(int2) * (size_t)fp.tilesX * fp.tilesY;
		for (size_t o = ((size_t)blockIdx.x * MR_GEOM_THREADS + tid) * 128; o < bytes; o += (size_t)gridDim.x * MR_GEOM_THREADS * 128)
			asm volatile("prefetch.global.L2 [%0];" ::"l"(tc + o));
	}
	// ---- statistics: one shared-memory atomic per warp, one RED per counter per CTA ----
	unsigned long long sum = acc;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		sum += __shfl_xor_sync(0xffffffffu, sum, o);
	if (lane == 0 && sum != 0ull)
		atomicAdd(&gs.stat, sum);
	__syncthreads();
	if (tid == 0)
	{
		const unsigned long long all = gs.stat;
		const int slot = blockIdx.x & (MR_STAT_SLOTS - 1);
		// (every warp of this CTA has seen the end of the work list, so every CTA has arrived and the list length is final)
		const unsigned listed = (blockIdx.x == 0) ? (unsigned)ldAcquire(syncTail) : 0u;
		if (listed + (unsigned)gs.ownCount)
			atomicAdd(&fp.ctr->visible, listed + (unsigned)gs.ownCount);
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			unsigned v = 0u;
			for (int w = 0; w < MR_GEOM_WARPS; w++)
				v = max(v, gs.dirty[w][k]);
			if (v)
				atomicMax(&fp.ctr->dirty[k], v);
		}
		if (blockIdx.x == 0)
		{
			fp.ctr->trianglesIn = (unsigned long long)fp.nTriReal;
			fp.ctr->clusters = (unsigned)nCl;
		}
		const unsigned long long nr = all & 0xfffffull, nc = (all >> 20) & 0xfffffull, nz = (all >> 40) & 0xfffffull;
		if (nr) atomicAdd(&fp.ctr->records[slot], nr);
		if (nc) atomicAdd(&fp.ctr->clippedIn[slot], nc);
		if (nz) atomicAdd(&fp.ctr->zeroCov[slot], nz);
	}
}

// ------------------------------------------------------------------------------------------
// Kernel 2 (only in frames with wide triangles): edge-chain checkpoints. The reference accumulates a row's edge
// functions column by column from the triangle's first column (e1 += n1.x, Renderer.cpp:243); a tile far to the right
// of that column would have to replay the whole prefix for each of its rows - quadratic in the triangle's width over
// its tiles (the X3D-style cloud scene of configs[3] has triangles 958 pixels wide: 184 of its 243 us went there).
// Here a warp takes one work item (a wide triangle, a block of 32 rows), a lane walks one row ONCE, with the
// reference's own additions in the reference's order, and stores (e1, e2) as they stand at every tile boundary; the
// tile kernel then starts at its own left edge with bit-identical values.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chain(const __grid_constant__ FrameParams fp)
{
	pdlLaunchDependents(); // k_raster's CTAs may become resident; they wait for this grid
	pdlWait();             // k_geom's records, work items and counters
	const int lane = threadIdx.x & 31;
	const int nW = (int)(gridDim.x * blockDim.x) >> 5, gw = (int)(blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int n = (int)min(fp.ctr->chkItems, (unsigned)fp.chkItemCap);
	for (int i = gw; i < n; i += nW)
	{
		const int4 it = fp.chkItems[i];
		if (it.x < 0)
			continue;
		const RecRef ref = recRef(fp, it.x);
		const F8 f01 = ldPair<L2_NORMAL>(ref.p), f23 = ldPair<L2_NORMAL>(ref.p + ref.stride), f89 = ldPair<L2_NORMAL>(ref.p + 4 * ref.stride);
		const uint32_t xspan = __float_as_uint(f23.b.x), yspan = __float_as_uint(f89.b.z);
		const int x0 = xspan & 0xffffu, x1 = xspan >> 16, y0 = yspan & 0xffffu, y1 = yspan >> 16;
		const int ntb = (x1 >> MR_TILE_SHIFT) - (x0 >> MR_TILE_SHIFT);
		{
			// binning of this block's tiles (k_geom leaves triangles with checkpoints to this kernel): a tile row belongs
			// to the block that holds its first row inside the bbox
			const int ra = y0 + it.z * 32, rb = min(ra + 31, y1);
			const int tx0 = x0 >> MR_TILE_SHIFT;
			for (int ty = max(ra >> MR_TILE_SHIFT, fp.tileRow0); ty <= min(rb >> MR_TILE_SHIFT, fp.tileRow0 + fp.tileRows - 1); ty++)
				if (max(ty << MR_TILE_SHIFT, y0) >= ra)
					for (int tx = tx0 + lane; tx <= tx0 + ntb; tx += 32)
					{
						const int tile = ty * fp.tilesX + tx;
						binStore(fp, tile, atomicAdd(&fp.tileCount[tile].x, 1), it.x);
					}
		}
		const int row = it.z * 32 + lane, y = y0 + row;
		if (y > y1 || y < fp.rowBegin || y >= fp.rowEnd)
			continue;
		const float n1x = f01.b.x, n1y = f01.b.y, n2x = f01.b.z, n2y = f01.b.w;
		const float ptx = (float)x0 + 0.5f, fy = (float)y + 0.5f;
		float e1 = n1x * (ptx - f01.a.z) + n1y * (fy - f01.a.w); // row start, Renderer.cpp:241-242
		float e2 = n2x * (ptx - f01.a.x) + n2y * (fy - f01.a.y);
		// entry (boundary tb, row) of a triangle's block is at tb * rows + row: the lanes of this warp (consecutive rows)
		// write, and the 16 rows of a tile read, consecutive entries
		const int rows = y1 - y0 + 1;
		float2* out = fp.chkPool + (size_t)it.y + (size_t)row;
		int x = x0;
		for (int tb = 0; tb < ntb; tb++)
		{
			const int xe = ((x0 >> MR_TILE_SHIFT) + tb + 1) << MR_TILE_SHIFT; // the next tile boundary
			for (; x < xe; x++)
			{
				e1 += n1x;
				e2 += n2x;
			}
			out[(size_t)tb * rows] = make_float2(e1, e2);
		}
	}
}

// ------------------------------------------------------------------------------------------
// Kernel 3: tile rasterizer + resolve + shader. One CTA per 16x16 tile.
// Phase 0: the tile's depth keys are fetched from gkeys (the small triangles' fragments have
//   already been depth-tested there by k_setup) and the entries are reset for the next frame.
// Phase 1 (only for tiles with binned triangles): coverage + depth of the larger triangles, with
//   the tile's keys in shared memory. A lane per binned triangle loads its record; each triangle is
//   then scanned by a quad of lanes (rows interleaved), reference loops D/E with the float edge
//   chain replayed from the triangle's own bbox start (e += n.x per column, Renderer.cpp:243).
//   Fragments go through a per-warp queue and are consumed 32 at a time by the whole warp: exact z,
//   then a 64-bit atomicMin in shared memory on (orderable z) << 32 | (record index + 1). The low
//   word makes equal-z fragments resolve to the earliest submitted triangle, which is what the
//   reference's strict `<` over in-order submission does.
// Phase 2 (thread per pixel): the winner's barycentrics are re-derived by the same chain, then
//   depth, perspective correction, texture and Blinn-Phong exactly as Renderer.cpp:253-305;
//   pixels without a winner get the clear values (Renderer.cpp:113-119) unless fp.keep.
// ------------------------------------------------------------------------------------------
#ifndef MR_RASTER_THREADS
#define MR_RASTER_THREADS 128
#endif
#define MR_FQ_CAP 64   // fragment queue entries per warp (power of two)
#define MR_FQ_SLOTS 64 // triangle slots per warp: 32 lanes per iteration, two iterations in flight

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
		const float z = pixelDepth(fp.persp, e1, e2, t.x, t.y, t.z);
		if (z == z) // a NaN depth never passes `z < pixdepth`
		{
			const unsigned long long key = ((unsigned long long)zkey(z) << 32) | (unsigned long long)__float_as_uint(t.w);
			unsigned long long* slot = &keys[info & 0xffu];
			if (key < *(volatile unsigned long long*)slot)
				atomicMin(slot, key);
		}
	}
}

// Appends this lane's fragment (if `on`) to the warp queue; consumes a batch when 32 are queued.
__device__ __forceinline__ void pushFragment(const FrameParams& fp, WarpQueue& wq, unsigned long long* keys, int& qhead, int& qcount,
                                             int lane, bool on, float e1, float e2, uint32_t info)
{
	const unsigned m = __ballot_sync(0xffffffffu, on);
	if (m == 0u)
		return;
	if (on)
	{
		const int w = (qhead + qcount + __popc(m & ((1u << lane) - 1u))) & (MR_FQ_CAP - 1);
		wq.e1[w] = e1;
		wq.e2[w] = e2;
		wq.info[w] = info;
	}
	qcount += __popc(m);
	if (qcount >= 32)
	{
		__syncwarp();
		consumeFragments(fp, wq, keys, qhead, 32, lane);
		__syncwarp(); // the consumed entries may be overwritten by the next pushes
		qhead = (qhead + 32) & (MR_FQ_CAP - 1);
		qcount -= 32;
	}
}

// ---- shading (Renderer.cpp:271-305) ----
// Coverage, depth and the winner of a pixel are bit-exact; its colour has to match the reference's within 1 LSB of the
// 8-bit image (north star), i.e. within 1/255 - seven orders of magnitude above float rounding. So everything
// downstream of the (exact) barycentrics uses the fast forms: fused multiply-adds (explicit intrinsics: the file is
// compiled with --fmad=false for the geometry), MUFU reciprocal / reciprocal square root instead of IEEE division and
// square root, and the specular power in float (integer exponents by squaring, others through exp2(y log2 x))
// instead of the reference's double pow(). Texture coordinates keep the reference's unfused expression: they select
// a texel. MR_EXACT_SHADING=1 restores the reference's operation order and double pow (float RGB bit-identical).

#if MR_EXACT_SHADING
__device__ __forceinline__ float interp3(float a, float b, float c, float k0, float k1, float k2) { return a * k0 + b * k1 + c * k2; }
__device__ __forceinline__ float sdot3(V3 a, V3 b) { return dot3(a, b); }
__device__ __forceinline__ float srlen3(V3 a) { return 1.0f / len3(a); }
__device__ __forceinline__ float sdiv(float a, float b) { return a / b; }
// pow(base, shininess) in double, as the reference evaluates it (its unqualified pow() is the double overload).
__device__ __forceinline__ float powShininess(float base, float shininess)
{
	const int n = (int)shininess;
	if ((float)n == shininess && n >= 1 && n <= 1024 && base >= 0.0f && base <= 2.0f)
	{
		double r = 1.0, b = (double)base;
		for (int e = n; e != 0; e >>= 1)
		{
			if (e & 1)
				r *= b;
			b *= b;
		}
		return (float)r;
	}
	return (float)pow((double)base, (double)shininess);
}
#else
__device__ __forceinline__ float interp3(float a, float b, float c, float k0, float k1, float k2) { return __fmaf_rn(c, k2, __fmaf_rn(b, k1, a * k0)); }
__device__ __forceinline__ float sdot3(V3 a, V3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float srlen3(V3 a) { return rsqrtf(sdot3(a, a)); }
__device__ __forceinline__ float sdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float powShininess(float base, float shininess)
{
	const int n = (int)shininess;
	if ((float)n == shininess && n >= 1 && n <= 64)
	{
		float r = 1.0f, b = base;
#pragma unroll
		for (int bit = 0; bit < 7; bit++) // warp-uniform exponent: a handful of multiplications
		{
			if ((n >> bit) & 1)
				r *= b;
			if ((n >> bit) > 1)
				b *= b;
		}
		return r;
	}
	return __powf(base, shininess); // exp2(y * log2(x)); base == 0 gives 0
}
#endif

// Renderer.cpp:271-305 for one pixel; returns the pixel value (and writes the normals image).
__device__ __forceinline__ V3 shadePixel(const FrameParams& fp, const MatDev& mat, float k0, float k1, float k2,
                                           const Corner& c0, const Corner& c1, const Corner& c2, const V3 ray, size_t pix)
{
	V3 color = mk3(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2]);
	V3 value = mk3(mat.emissive[0], mat.emissive[1], mat.emissive[2]);
	if (fp.texturing && mat.texOffset >= 0 && mat.texRows > 0)
	{
		const float u = c0.u * k0 + c1.u * k1 + c2.u * k2;
		const float v = c0.v * k0 + c1.v * k1 + c2.v * k2;
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
		const V3 position = fp.unproject ? ray // (Renderer.cpp:281 up to rounding: the point of the triangle's plane on this pixel's ray)
		                                 : mk3(interp3(c0.px, c1.px, c2.px, k0, k1, k2), interp3(c0.py, c1.py, c2.py, k0, k1, k2), interp3(c0.pz, c1.pz, c2.pz, k0, k1, k2));
		const V3 light = mk3(fp.light[0], fp.light[1], fp.light[2]);
		V3 lightdir = light;
		if (fp.lightIsPoint)
		{
			const V3 l = sub3(light, position);
			lightdir = scale3(l, srlen3(l));
		}
		const V3 normal = mk3(interp3(c0.nx, c1.nx, c2.nx, k0, k1, k2), interp3(c0.ny, c1.ny, c2.ny, k0, k1, k2), interp3(c0.nz, c1.nz, c2.nz, k0, k1, k2));
		const float nl = sdot3(normal, lightdir);
		const float rnlen = srlen3(normal); // 1 / |normal|
#if MR_EXACT_SHADING
		const float nlen = len3(normal);
		const float d = ((0.0f > nl) ? 0.0f : nl) / nlen + fp.ambient;
		value = add3(value, scale3(color, d));
#else
		const float d = __fmaf_rn((0.0f > nl) ? 0.0f : nl, rnlen, fp.ambient);
		value = mk3(__fmaf_rn(color.x, d, value.x), __fmaf_rn(color.y, d, value.y), __fmaf_rn(color.z, d, value.z));
#endif
		if (mat.shininess != 0.0f)
		{
			const V3 viewdir = scale3(position, srlen3(position));
			const V3 hv = sub3(lightdir, viewdir);
			const float hn = sdot3(hv, normal);
#if MR_EXACT_SHADING
			const float base = ((hn > 0.0f) ? hn : 0.0f) / (len3(hv) * nlen);
			const float specular = powShininess(base, mat.shininess);
			value = add3(value, scale3(mk3(mat.specular[0], mat.specular[1], mat.specular[2]), specular));
#else
			const float base = ((hn > 0.0f) ? hn : 0.0f) * (srlen3(hv) * rnlen);
			const float specular = powShininess(base, mat.shininess);
			value = mk3(__fmaf_rn(mat.specular[0], specular, value.x), __fmaf_rn(mat.specular[1], specular, value.y), __fmaf_rn(mat.specular[2], specular, value.z));
#endif
		}
		if (fp.saveNormals && fp.normals)
		{
			float* pn = fp.normals + 3 * pix;
			pn[0] = normal.x; pn[1] = normal.y; pn[2] = normal.z;
		}
	}
	return value;
}

// Tile output staging: 16 rows x (48 rgb floats + 16 depth floats); written to HBM as 32-byte sectors
// (full tiles) or float4 rows.
struct TileOut
{
	// per pixel row: 48 rgb floats then 16 depth floats = 8 sectors of 32 bytes, in the order the tile store
	// sends them out; rows padded to 80 floats: 32-byte aligned, and 80 mod 32 = 16 puts the 8-byte staging
	// stores of two neighbouring rows (a half warp) on disjoint banks
	float px[MR_TILE][80];
};

// Writes a tile's rows to the framebuffer as float4: 12 per row of rgb (192 B), 4 per row of depth
// (64 B); every store instruction covers whole 32-byte sectors. `to` == 0 writes the clear values.
__device__ __forceinline__ void storeTileRows(const FrameParams& fp, int tileX0, int tileY0, int tid, const TileOut* to)
{
	if (tid < MR_TILE * 12)
	{
		const int row = tid / 12, j = tid - row * 12, y = tileY0 + row;
		if (y < fp.h && y >= fp.rowBegin && y < fp.rowEnd)
		{
			float4 v;
			if (to)
				v = *reinterpret_cast<const float4*>(&to->px[row][4 * j]);
			else
			{
				// channel of the first of the four floats: (4 j) mod 3 == j mod 3
				const int c = j % 3;
				const float r = fp.bg[0], g = fp.bg[1], b = fp.bg[2];
				const float b0 = (c == 0) ? r : (c == 1) ? g : b, b1 = (c == 0) ? g : (c == 1) ? b : r, b2 = (c == 0) ? b : (c == 1) ? r : g;
				v = make_float4(b0, b1, b2, b0);
			}
			*reinterpret_cast<float4*>(fp.image + 3 * ((size_t)y * fp.w + tileX0) + 4 * j) = v;
		}
	}
	else if (tid < MR_TILE * 12 + MR_TILE * 4)
	{
		const int t = tid - MR_TILE * 12, row = t >> 2, j = t & 3, y = tileY0 + row;
		if (y < fp.h && y >= fp.rowBegin && y < fp.rowEnd)
		{
			const float4 v = to ? *reinterpret_cast<const float4*>(&to->px[row][48 + 4 * j]) : make_float4(1e11f, 1e11f, 1e11f, 1e11f);
			*reinterpret_cast<float4*>(fp.depth + (size_t)y * fp.w + tileX0 + 4 * j) = v;
		}
	}
}

// The same for a tile that lies fully inside the image and the strip (almost every tile), by a CTA of
// 128 threads: a tile is 16 rows x (6 sectors of rgb + 2 sectors of depth) = 128 sectors of 32 bytes, one
// 256-bit store (STG.256) per thread, no bounds tests. Needs an image width that is a multiple of 8.
template <bool CLEAR>
__device__ __forceinline__ void storeFullTile128(const FrameParams& fp, int tileX0, int tileY0, int tid, const TileOut* to)
{
	const int row = tid >> 3, j = tid & 7;
	const size_t pix = (size_t)(tileY0 + row) * fp.w + tileX0;
	float* dst = (j < 6) ? fp.image + 3 * pix + 8 * j : fp.depth + pix + 8 * (j - 6);
	float4 v0, v1;
	if (CLEAR)
	{
		// rgb sector j starts at float 8 j of the row: channel (8 j) mod 3 = (2 j) mod 3; bgPattern = r g b r g b ...
		const int c = (2 * j) % 3;
		v0 = (j < 6) ? make_float4(fp.bgPattern[c], fp.bgPattern[c + 1], fp.bgPattern[c + 2], fp.bgPattern[c + 3]) : make_float4(1e11f, 1e11f, 1e11f, 1e11f);
		v1 = (j < 6) ? make_float4(fp.bgPattern[c + 4], fp.bgPattern[c + 5], fp.bgPattern[c + 6], fp.bgPattern[c + 7]) : v0;
	}
	else
	{
		v0 = *reinterpret_cast<const float4*>(&to->px[row][8 * j]);
		v1 = *reinterpret_cast<const float4*>(&to->px[row][8 * j + 4]);
	}
	stPair<true, MR_TILE_STORE_HINT>(reinterpret_cast<F8*>(dst), v0, v1);
}

// Phase 1 for one batch of up to 32 binned triangles (a lane each; `have` lanes hold record `id`).
__device__ __forceinline__ void rasterBatch(const FrameParams& fp, WarpQueue& wq, unsigned long long* keys, int& qhead, int& qcount,
                                            int& parity, int lane, bool have, int id, int tileX0, int tileY0)
{
	const int slot = parity * 32 + lane;
	float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
	if (have)
	{
		const RecRef ref = recRef(fp, id);
		// (binned triangles always carry pair 4: rows and flags)
		const F8 f01 = ldPair(ref.p), f23 = ldPair(ref.p + ref.stride), f89 = ldPair(ref.p + 4 * ref.stride);
		q0 = f01.a; q1 = f01.b; q2 = f23.a;
		q3 = make_float4(f23.b.x, f89.b.z, f89.b.w, 0.0f); // column span, row span, flags
		wq.tri[slot] = make_float4(q2.x, q2.y, q2.z, __uint_as_float((uint32_t)(id + 1)));
	}
	const uint32_t xspan = __float_as_uint(q3.x), yspan = __float_as_uint(q3.y);
	const int x0 = xspan & 0xffffu, x1 = xspan >> 16, y0 = yspan & 0xffffu, y1 = yspan >> 16;
	__syncwarp();

	// a group of lanes scans each triangle of the batch, rows interleaved: quads while the warp holds more than four
	// triangles, 8 or 16 lanes each for the last few (a tile with a handful of big triangles keeps every lane busy)
	unsigned large = __ballot_sync(0xffffffffu, have);
	while (large != 0u)
	{
		const int left = __popc(large);
		const int gshift = (left <= 2) ? 4 : (left <= 4) ? 3 : 2, G = 1 << gshift;
		const int q = lane & (G - 1), quad = lane >> gshift;
		const int src = __fns(large, 0, quad + 1); // lane that owns this group's triangle, or -1
		const bool on = src >= 0 && src < 32;
		const int s = on ? src : 0;
		const float p0x = __shfl_sync(0xffffffffu, q0.x, s), p0y = __shfl_sync(0xffffffffu, q0.y, s);
		const float p2x = __shfl_sync(0xffffffffu, q0.z, s), p2y = __shfl_sync(0xffffffffu, q0.w, s);
		const float n1x = __shfl_sync(0xffffffffu, q1.x, s), n1y = __shfl_sync(0xffffffffu, q1.y, s);
		const float n2x = __shfl_sync(0xffffffffu, q1.z, s), n2y = __shfl_sync(0xffffffffu, q1.w, s);
		const int ry0 = __shfl_sync(0xffffffffu, y0, s);
		const int bx0 = __shfl_sync(0xffffffffu, x0, s), bx1 = min(__shfl_sync(0xffffffffu, x1, s), tileX0 + MR_TILE - 1);
		const int ry1 = __shfl_sync(0xffffffffu, y1, s);
		const int by0 = max(ry0, tileY0), by1 = min(ry1, tileY0 + MR_TILE - 1);
		// checkpoints of a wide triangle's chain (k_chain): 1 + its first pool entry, 0 = none
		const uint32_t chk = __shfl_sync(0xffffffffu, __float_as_uint(q3.z), s) >> 1;
		const int tslot = parity * 32 + s;
		const int xs = max(bx0, tileX0);
		const int ncols = on ? max(bx1 - xs + 1, 0) : 0;
		const int myRows = on ? max((by1 - by0 - q + G) >> gshift, 0) : 0; // rows by0+q, by0+q+G, ...
		const float ptx = (float)bx0 + 0.5f;
		const int maxRows = __reduce_max_sync(0xffffffffu, myRows);
		const int maxCols = __reduce_max_sync(0xffffffffu, ncols);
		for (int rr = 0; rr < maxRows; rr++)
		{
			const bool rowOn = rr < myRows;
			const int y = by0 + q + G * rr;
			const float fy = (float)y + 0.5f;
			float e1 = n1x * (ptx - p2x) + n1y * (fy - p2y);
			float e2 = n2x * (ptx - p0x) + n2y * (fy - p0y);
			if (rowOn && xs > bx0)
			{
				if (chk != 0u)
				{
					// the chain as it stands at this tile's left edge, accumulated once by k_chain
					const float2 cp = __ldg(fp.chkPool + (size_t)(chk - 1u) + (size_t)((tileX0 >> MR_TILE_SHIFT) - (bx0 >> MR_TILE_SHIFT) - 1) * (ry1 - ry0 + 1) + (y - ry0));
					e1 = cp.x;
					e2 = cp.y;
				}
				else
					for (int x = bx0; x < xs; x++) // chain prefix left of the tile
					{
						e1 += n1x;
						e2 += n2x;
					}
			}
			const uint32_t rowInfo = (uint32_t)((y - tileY0) * MR_TILE + (xs - tileX0)) | ((uint32_t)tslot << 8);
			for (int cc = 0; cc < maxCols; cc++, e1 += n1x, e2 += n2x)
			{
				const float k0 = 1.0f - e1 - e2;
				const bool inside = rowOn && cc < ncols && insideTest(e1, e2, k0); // Renderer.cpp:245
				pushFragment(fp, wq, keys, qhead, qcount, lane, inside, e1, e2, rowInfo + (uint32_t)cc);
			}
		}
		// drop the triangles just done
		for (int k = 0; k < (32 >> gshift) && large != 0u; k++)
			large &= large - 1u;
	}
	// Triangle slots of this parity are overwritten two batches from now; fragments still queued
	// then must not refer to them, so drain before reusing a parity.
	if (parity == 1 && qcount > 0)
	{
		__syncwarp();
		consumeFragments(fp, wq, keys, qhead, qcount, lane);
		__syncwarp();
		qhead = (qhead + qcount) & (MR_FQ_CAP - 1);
		qcount = 0;
	}
	parity ^= 1;
	__syncwarp();
}

// Phase 1 of a tile with binned triangles (see k_raster): their coverage and depth into the tile's shared-memory keys.
// (ptxas 12.9 crashes on this as a __noinline__ function; inlined, it sits between the hot parts of k_raster without being fetched. The
// issue rate of a kernel falls off once its hot instructions exceed the ~32 KB instruction cache: tools/ubench/icache.cu.)
__device__ __forceinline__ void rasterBinned(const FrameParams& fp, int tile, int total, unsigned long long ovfTotal, unsigned long long* keys, WarpQueue* queues,
                                          int tileX0, int tileY0)
{
	const int tid = threadIdx.x, lane = tid & 31;
	WarpQueue& wq = queues[tid >> 5];
	const int count = min(total, fp.binCap);
	const int* bin = fp.bins + (size_t)tile * fp.binCap;
	int qhead = 0, qcount = 0; // warp-uniform
	int parity = 0;
	// (the bin's triangles are dealt round-robin to the CTA's warps: a short bin is shared by all of them)
	for (int base = 0; base < count; base += MR_RASTER_THREADS)
	{
		const int i = base + lane * (MR_RASTER_THREADS / 32) + (tid >> 5);
		const bool have = i < count;
		const int id = have ? __ldg(&bin[i]) : 0;
		rasterBatch(fp, wq, keys, qhead, qcount, parity, lane, have, id, tileX0, tileY0);
	}
	if (total > fp.binCap)
	{
		// this tile spilled: its remaining triangles are somewhere in the global overflow list
		const unsigned long long n = min(ovfTotal, (unsigned long long)fp.ovfCap);
		for (unsigned long long base = (unsigned long long)(tid >> 5) * 32; base < n; base += MR_RASTER_THREADS)
		{
			const unsigned long long i = base + lane;
			int2 p = make_int2(-1, 0);
			if (i < n)
				p = __ldg(&fp.ovfPairs[i]);
			const bool have = p.x == tile;
			if (__any_sync(0xffffffffu, have))
				rasterBatch(fp, wq, keys, qhead, qcount, parity, lane, have, p.y, tileX0, tileY0);
		}
	}
	if (qcount > 0)
	{
		__syncwarp();
		consumeFragments(fp, wq, keys, qhead, qcount, lane);
	}
}

#ifndef MR_PREFIX_SHARE_MIN
#define MR_PREFIX_SHARE_MIN 8
#endif
// Per-pixel part of phase 2 for tile pixel `pi` (0..255) whose final depth key is `key`.
// Warp-convergent: contains shuffles. Returns the pixel's colour and depth (clear values when
// nothing won); side outputs (winner ids, normals image) are written here.
template <int TM>
__device__ __forceinline__ void resolvePixel(const FrameParams& fp, unsigned long long key, int pi, int tileX0, int tileY0, int lane,
                                             V3& value, float& zout, bool& store)
{
	const int px = tileX0 + (pi & 15), py = tileY0 + (pi >> 4);
	const bool inImage = px < fp.w && py < fp.h && py >= fp.rowBegin && py < fp.rowEnd;
	const size_t pix = (size_t)py * fp.w + px;
	const uint32_t win = inImage ? (uint32_t)(key & 0xffffffffull) : 0u;
	const bool attrs = fp.lighting || fp.texturing; // does shading read the corners at all?
	float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
	float4 s0 = q0, s1 = q0, s2 = q0, s3 = q0, s4 = q0, s5 = q0;
	int id = -1;
	if (win != 0u)
	{
		id = (int)(win - 1u);
		const RecRef ref = recRef(fp, id);
		const F8* r8 = ref.p;
		const int st = ref.stride;
		const F8 f01 = ldPair(r8), f23 = ldPair(r8 + st);
		q0 = f01.a; q1 = f01.b; q2 = f23.a; q3 = f23.b; // q3 = (column span, nx of the three corners)
		if (attrs || fp.winner)
		{
			const F8 f45 = ldPair(r8 + 2 * st); // the other normal components, submission index
			s0 = f45.a; s1 = f45.b;
		}
		if (fp.lighting && !fp.unproject)
		{
			const F8 f67 = ldPair(r8 + 3 * st); // corner positions
			s2 = f67.a; s3 = f67.b;
		}
		if (fp.texturing) // texture coordinates; rows and checkpoints of wide triangles (the chain is only replayed for texel selection)
		{
			const F8 f89 = ldPair(r8 + 4 * st);
			s4 = f89.a; s5 = f89.b;
		}
	}
	// Replay of the winner's edge chain. Lanes of the same pixel row (a warp covers 2 or 4 rows) that share a
	// winner starting left of the tile share the prefix of the chain up to the tile edge: one
	// lane walks it, the others receive it by shuffle and only add their in-tile columns.
	// Only texel selection has to follow the reference's ACCUMULATED edge functions (a texture coordinate a few 1e-7 off can
	// pick the neighbouring texel); without texturing the barycentrics only weight the shading inputs, so they are
	// evaluated directly at the pixel centre (the row-start expression of Renderer.cpp:241-242 with x at this pixel),
	// and the depth - which must be exact - is the winner's own, taken out of the depth key.
	const bool replay = fp.texturing != 0;
	const int x0 = replay ? (int)(__float_as_uint(q3.x) & 0xffffu) : px;
	const float ptx = (float)x0 + 0.5f, fy = (float)py + 0.5f;
	float e1 = q1.x * (ptx - q0.z) + q1.y * (fy - q0.w);
	float e2 = q1.z * (ptx - q0.x) + q1.w * (fy - q0.y);
	int xcur = x0;
	if (replay)
	{
		// (a prefix of a few columns is cheaper walked by every lane than matched and shuffled)
		int prefix = (id >= 0) ? tileX0 - x0 : 0; // columns left of the tile
		const uint32_t chk = fp.chkEnable ? __float_as_uint(s5.w) >> 1 : 0u;
		if (prefix > 0 && chk != 0u)
		{
			// a wide triangle with checkpoints (k_chain): the chain as it stands at this tile's left edge
			const int ry0 = (int)(__float_as_uint(s5.z) & 0xffffu), ry1 = (int)(__float_as_uint(s5.z) >> 16);
			const float2 cp = __ldg(fp.chkPool + (size_t)(chk - 1u) + (size_t)((tileX0 >> MR_TILE_SHIFT) - (x0 >> MR_TILE_SHIFT) - 1) * (ry1 - ry0 + 1) + (py - ry0));
			e1 = cp.x;
			e2 = cp.y;
			xcur = tileX0;
			prefix = 0;
		}
		if (prefix <= MR_PREFIX_SHARE_MIN)
			prefix = 0;
		if (__any_sync(0xffffffffu, prefix > 0))
		{
			const unsigned long long groupKey = (prefix > 0) ? (((unsigned long long)(uint32_t)id << 2) | (unsigned long long)((pi >> 4) & 3))
			                                                 : (0x8000000000000000ull | (unsigned long long)lane);
			const unsigned peers = __match_any_sync(0xffffffffu, groupKey);
			const int leader = __ffs(peers) - 1;
			if (lane == leader && prefix > 0)
				for (int x = 0; x < prefix; x++)
				{
					e1 += q1.x;
					e2 += q1.z;
				}
			const float t1 = __shfl_sync(0xffffffffu, e1, leader), t2 = __shfl_sync(0xffffffffu, e2, leader);
			if (prefix > 0)
			{
				e1 = t1;
				e2 = t2;
				xcur = tileX0;
			}
		}
	}
	// per-pixel result: the clear values unless a triangle won the pixel
	value = mk3(fp.bg[0], fp.bg[1], fp.bg[2]);
	zout = 1e11f;
	const bool won = inImage && win != 0u;
	store = inImage && (won || !fp.keep);
	if (won)
	{
		for (int x = xcur; x < px; x++)
		{
			e1 += q1.x;
			e2 += q1.z;
		}
		float k0 = 1.0f - e1 - e2, k1 = e1, k2 = e2;
		// the winning fragment's depth as k_geom / phase 1 computed it (Renderer.cpp:255 / :261 on the accumulated edge
		// functions), bit for bit, from the key; a key cannot tell -0 from +0, so a zero is computed again
		const float zk = unzkey((uint32_t)(key >> 32));
		if (fp.persp)
		{
			zout = (zk != 0.0f) ? zk : 1.0f / (k0 * q2.x + k1 * q2.y + k2 * q2.z);
			k0 *= q2.x * zout;
			k1 *= q2.y * zout;
			k2 *= q2.z * zout;
		}
		else
			zout = (zk != 0.0f) ? zk : k0 * q2.x + k1 * q2.y + k2 * q2.z + 0.0f * 1.0f;
		if (fp.winner)
			fp.winner[pix] = (int)__float_as_uint(s1.z); // the reference's submission index (instance ids are padded per renderable)
		const MatDev& mat = frameMats<TM>(fp)[__float_as_uint(q2.w)];
		Corner c0, c1, c2;
		// (layout: storeRec)
		c0.nx = q3.y; c1.nx = q3.z; c2.nx = q3.w;
		c0.ny = s0.x; c0.nz = s0.y; c1.ny = s0.z; c1.nz = s0.w;
		c2.ny = s1.x; c2.nz = s1.y;
		c0.px = s2.x; c0.py = s2.y; c0.pz = s2.z;
		c1.px = s2.w; c1.py = s3.x; c1.pz = s3.y;
		c2.px = s3.z; c2.py = s3.w; c2.pz = s1.w;
		c0.u = s4.x; c0.v = s4.y; c1.u = s4.z; c1.v = s4.w;
		c2.u = s5.x; c2.v = s5.y;
		// under the standard perspective the pixel's view-space position is its own ray scaled by the depth
		const V3 ray = mk3(__fmaf_rn(fp.unprojX[0], (float)px + 0.5f, fp.unprojX[1]) * zout, __fmaf_rn(fp.unprojY[0], (float)py + 0.5f, fp.unprojY[1]) * zout, -zout);
		value = shadePixel(fp, mat, k0, k1, k2, c0, c1, c2, ray, pix);
	}
	else if (inImage && !fp.keep)
	{
		if (fp.saveNormals && fp.normals)
		{
			float* pn = fp.normals + 3 * pix;
			pn[0] = 0.0f; pn[1] = 0.0f; pn[2] = 1.0f;
		}
		if (fp.winner)
			fp.winner[pix] = -1;
	}
}

// One tile: phases 0, 1 and 2, by a CTA of NT threads (128: each thread resolves two pixels; twice
// as many tiles are then in flight per SM). Called by all threads of the CTA (contains barriers).
// With two pixels per thread they are neighbours in a row (tile pixels 2 tid, 2 tid + 1): both depth keys
// come in one 16-byte load, and the thread's part of the staged tile is 24 + 8 contiguous bytes.
template <int NT>
__device__ __forceinline__ int tilePixel(int tid, int pp) { return (MR_TILE_PIXELS / NT == 2) ? 2 * tid + pp : tid + pp * NT; }

template <int NT, int TM>
__device__ __forceinline__ void rasterTile(const FrameParams& fp, int tx, int ty, unsigned long long* keys, WarpQueue* queues)
{
	constexpr int PP = MR_TILE_PIXELS / NT;
	const int tile = ty * fp.tilesX + tx;
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int tileX0 = tx * MR_TILE, tileY0 = ty * MR_TILE;
	// float4 row stores need 16-byte aligned rows and a tile that lies fully inside the image width
	const bool vec = ((fp.w & 3) == 0) && (tileX0 + MR_TILE <= fp.w) && !fp.keep;
	// the common case: every pixel of the tile belongs to this frame (no per-pixel bounds tests)
	const bool full = PP == 2 && NT == 128 && tx < fp.fullTx && ty >= fp.fullTy0 && ty < fp.fullTy1;
	// x: larger triangles binned to this tile, y: fragments of small triangles may have reached its keys
	const int2 tinfo = fp.tileCount[tile];
	const int total = tinfo.x;
	if (full && tinfo.x == 0 && tinfo.y == 0 && !(fp.saveNormals && fp.normals) && !fp.winner)
	{
		// nothing has touched this tile: clear values only (Renderer.cpp:113-119), its keys are not even read
		if (!fp.sparseStores)
			storeFullTile128<true>(fp, tileX0, tileY0, tid, 0);
		return;
	}
	// ---- phase 0: the tile's keys out of gkeys, merged with the clear / kept depth; gkeys reset ----
	unsigned long long key[PP];
	const unsigned long long clearKey = (unsigned long long)zkey(1e11f) << 32;
	if (full)
	{
		const int pi = 2 * tid;
		ulonglong2* gp = reinterpret_cast<ulonglong2*>(fp.gkeys + (size_t)(tileY0 + (pi >> 4)) * fp.w + tileX0 + (pi & 15));
		const ulonglong2 g = __ldcs(gp); // read once per frame: streaming
		if ((g.x & g.y) != MR_KEY_EMPTY)
			*gp = make_ulonglong2(MR_KEY_EMPTY, MR_KEY_EMPTY); // ready for the next frame
		key[0] = (g.x < clearKey) ? g.x : clearKey;
		key[PP - 1] = (g.y < clearKey) ? g.y : clearKey;
	}
	else
	{
#pragma unroll
		for (int pp = 0; pp < PP; pp++)
		{
			const int pi = tilePixel<NT>(tid, pp);
			const int px = tileX0 + (pi & 15), py = tileY0 + (pi >> 4);
			unsigned long long g = MR_KEY_EMPTY, base = 0ull; // pixels outside the image / strip can never be won
			if (px < fp.w && py < fp.h && py >= fp.rowBegin && py < fp.rowEnd)
			{
				const size_t pix = (size_t)py * fp.w + px;
				g = __ldcs(&fp.gkeys[pix]);
				base = fp.keep ? (unsigned long long)zkey(fp.depth[pix]) << 32 : clearKey;
				if (g != MR_KEY_EMPTY)
					fp.gkeys[pix] = MR_KEY_EMPTY;
			}
			key[pp] = (g < base) ? g : base;
		}
	}
	unsigned long long ovfTotal = 0ull;
	if (total > fp.binCap)
	{
		// this tile spilled into the overflow list (rare)
		ovfTotal = __ldg(&fp.ctr->ovfTotal);
		if (threadIdx.x == 0)
			atomicMax(&fp.ctr->maxTile, (unsigned)total); // lets the host size the bins for the next frames
		// if the overflow list itself overflowed, the host regrows it, clears gkeys and re-runs the frame
	}

	// ---- phase 1: coverage + depth of the binned triangles ----
	if (total > 0)
	{
#pragma unroll
		for (int pp = 0; pp < PP; pp++)
			keys[tilePixel<NT>(tid, pp)] = key[pp];
		if (tid == 0)
			atomicAdd(&fp.ctr->pairTotal[tile & (MR_STAT_SLOTS - 1)], (unsigned long long)total);
		__syncthreads();
		rasterBinned(fp, tile, total, ovfTotal, keys, queues, tileX0, tileY0);
		__syncthreads();
#pragma unroll
		for (int pp = 0; pp < PP; pp++)
			key[pp] = keys[tilePixel<NT>(tid, pp)];
	}

	// ---- empty tile: clear values only (Renderer.cpp:113-119) ----
	bool anyWin = false;
#pragma unroll
	for (int pp = 0; pp < PP; pp++)
		anyWin = anyWin || (uint32_t)(key[pp] & 0xffffffffull) != 0u;
	// (this barrier also separates the previous tile's staging reads from this tile's staging writes)
	const bool anyWinTile = __syncthreads_or(anyWin);
	if (tid == 0 && (tinfo.x | tinfo.y) != 0)
	{
		fp.tileCount[tile] = make_int2(0, 0); // every thread has read it: ready for the next frame
		atomicAdd(&fp.ctr->tilesStored[tile & (MR_STAT_SLOTS - 1)], 1ull);
	}
	if (!anyWinTile && vec && !(fp.saveNormals && fp.normals) && !fp.winner)
	{
		if (full)
			storeFullTile128<true>(fp, tileX0, tileY0, tid, 0);
		else
			for (int pi = tid; pi < MR_TILE_PIXELS; pi += NT)
				storeTileRows(fp, tileX0, tileY0, pi, 0);
		return;
	}

	// ---- phase 2: resolve + shade, 256 / NT pixels per thread, one after the other through the same code (a second
	// inlined copy would double the hot instructions); results are staged in shared memory (the fragment queues are
	// idle now) and written as whole rows ----
	TileOut* to = reinterpret_cast<TileOut*>(queues);
#pragma unroll 1
	for (int pp = 0; pp < PP; pp++)
	{
		V3 value;
		float zout;
		bool store;
		const int pi = tilePixel<NT>(tid, pp);
		resolvePixel<TM>(fp, (pp == 0) ? key[0] : key[PP - 1], pi, tileX0, tileY0, lane, value, zout, store);
		if (vec)
		{
			const int r = pi >> 4, c = pi & 15;
			to->px[r][3 * c] = value.x;
			to->px[r][3 * c + 1] = value.y;
			to->px[r][3 * c + 2] = value.z;
			to->px[r][48 + c] = zout;
		}
		else if (store)
		{
			const size_t pix = (size_t)(tileY0 + (pi >> 4)) * fp.w + tileX0 + (pi & 15);
			float* img = fp.image + 3 * pix;
			img[0] = value.x; img[1] = value.y; img[2] = value.z;
			fp.depth[pix] = zout;
		}
	}
	if (vec)
	{
		__syncthreads();
		if (full)
			storeFullTile128<false>(fp, tileX0, tileY0, tid, to);
		else
			for (int pi = tid; pi < MR_TILE_PIXELS; pi += NT)
				storeTileRows(fp, tileX0, tileY0, pi, to);
	}
}

// One CTA per tile. (Persistent CTAs measured slower twice: with a global tile counter, -10 %, and
// with a static round-robin plus next-tile prefetch, 46 vs 44 us on the sphere and 255 vs 189 us on
// the cloud scene — the hardware CTA scheduler balances 8160 uneven tiles better.)
#ifndef MR_RASTER_MINB
#define MR_RASTER_MINB (1024 / MR_RASTER_THREADS)
#endif
template <int TM>
__global__ void __launch_bounds__(MR_RASTER_THREADS, MR_RASTER_MINB) k_raster(const __grid_constant__ FrameParams fp)
{
	__shared__ unsigned long long keys[MR_TILE_PIXELS];
	__shared__ WarpQueue queues[MR_RASTER_THREADS / 32 < 4 ? 4 : MR_RASTER_THREADS / 32]; // also >= sizeof(TileOut)
#ifdef MR_TIMELINE
	const int tlTile = (blockIdx.y * gridDim.x + blockIdx.x);
	if (threadIdx.x == 0 && tlTile < 8192) g_rtl[tlTile * 4 + 0] = globalTimer();
#endif
	pdlWait(); // k_geom's keys, bins and records
#ifdef MR_TIMELINE
	if (threadIdx.x == 0 && tlTile < 8192) g_rtl[tlTile * 4 + 1] = globalTimer();
#endif
	if (blockIdx.x == 0 && blockIdx.y == 0)
	{
		// k_geom is done with its work list and this frame's statistics are where they belong:
		// the list counters and the next frame's statistics start from zero
		if (threadIdx.x < 3)
			fp.geomSync[threadIdx.x * MR_SYNC_STRIDE] = 0;
		for (int i = threadIdx.x; i < (int)(sizeof(Counters) / 8); i += MR_RASTER_THREADS)
			reinterpret_cast<unsigned long long*>(fp.ctrNext)[i] = 0ull;
	}
	rasterTile<MR_RASTER_THREADS, TM>(fp, blockIdx.x, fp.tileRow0 + blockIdx.y, keys, queues);
#ifdef MR_TIMELINE
	if (threadIdx.x == 0 && tlTile < 8192)
	{
		unsigned smid;
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		g_rtl[tlTile * 4 + 2] = globalTimer();
		g_rtl[tlTile * 4 + 3] = smid;
	}
#endif
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

// Reads a buffer larger than the L2 (second half of mr_flush_l2): after the memset the L2 is
// full of *dirty* flush lines, whose write-back would otherwise be charged to the next kernels;
// streaming reads replace them with clean lines.
__global__ void k_flush_read(const float4* __restrict__ src, size_t n, float* sink)
{
	float acc = 0.0f;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
	{
		const float4 v = __ldcs(&src[i]);
		acc += v.x + v.y + v.z + v.w;
	}
	if (acc == 12345.678f)
		*sink = acc;
}

// Clear values into the pixels [p0, p1) of an image / depth pair (mr_clear_rows).
__global__ void k_clear_rows(float* image, float* depth, size_t p0, size_t p1, float r, float g, float b)
{
	const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (size_t p = p0 + first; p < p1; p += stride)
	{
		image[3 * p] = r;
		image[3 * p + 1] = g;
		image[3 * p + 2] = b;
		depth[p] = 1e11f;
	}
}

// Stream-ordered flags in (peer) device memory: see mr_stream_signal / mr_stream_wait.
__global__ void k_signal(unsigned* word, unsigned value)
{
	__threadfence_system(); // everything this stream did before is visible to the other GPUs first
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(word), "r"(value) : "memory");
}

__global__ void k_wait(const unsigned* words, int n, unsigned value)
{
	for (int i = threadIdx.x; i < n; i += blockDim.x)
	{
		unsigned v;
		do
		{
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(words + i) : "memory");
		} while ((int)(v - value) < 0);
	}
}

__global__ void k_selftest(const float* in, float* out)
{
	// with contraction, a*b+c keeps the exact product; without, the product rounds first
	out[0] = in[0] * in[1] + in[2];
}

}

// Shape of k_geom for meshlets of up to nvCap corners: dynamic shared memory per CTA and a grid of as many
// CTAs as are resident at once (they are persistent: work comes from the frame's slice counter).
int mrk_geom_config(int nvCap, int smCount, int* grid, int* smemBytes)
{
	int dev = 0, smemMax = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
		return -1;
	const int bytes = MR_GEOM_FIXED_BYTES + MR_GEOM_WARPS * MR_GEOM_WARP_BYTES(nvCap);
	if (bytes > smemMax)
		return -1;
	int perSm = 0;
	for (int tm = 0; tm < 2; tm++)
	{
		const void* fn = tm ? (const void*)k_geom<TM_INLINE> : (const void*)k_geom<TM_GLOBAL>;
		if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
			return -1;
		int n = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, MR_GEOM_THREADS, bytes) != cudaSuccess || n < 1)
			return -1;
		perSm = tm ? std::min(perSm, n) : n;
	}
	*grid = smCount * std::min(perSm, MR_GEOM_MINB);
	*smemBytes = bytes;
	return 0;
}

void mrk_launch_frame(const FrameParams& fp, int geomGrid, int geomSmem, cudaStream_t stream, cudaEvent_t* ev, cudaEvent_t bracketStart, cudaEvent_t bracketStop, bool pdl,
                      const unsigned* gateWord, unsigned gateValue)
{
	if (bracketStart) cudaEventRecord(bracketStart, stream);
	if (ev) cudaEventRecord(ev[0], stream);
	const bool inl = fp.inlineTables != 0;
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	cfg.attrs = attr;
	{
		// all CTAs of k_geom must be resident (grid-wide rendezvous after the cull phase): cooperative launch
		attr[0].id = cudaLaunchAttributeCooperative;
		attr[0].val.cooperative = 1;
		cfg.numAttrs = 1;
		cfg.gridDim = dim3(std::max(1, geomGrid));
		cfg.blockDim = dim3(MR_GEOM_THREADS);
		cfg.dynamicSmemBytes = (size_t)geomSmem;
		if (inl)
			cudaLaunchKernelEx(&cfg, k_geom<TM_INLINE>, fp);
		else
			cudaLaunchKernelEx(&cfg, k_geom<TM_GLOBAL>, fp);
	}
	if (fp.chkEnable && fp.tileRows > 0)
	{
		// frames with wide triangles: the checkpoints of their edge chains, between the two kernels
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.numAttrs = pdl ? 1 : 0;
		cfg.gridDim = dim3(148 * 4);
		cfg.blockDim = dim3(128);
		cfg.dynamicSmemBytes = 0;
		cudaLaunchKernelEx(&cfg, k_chain, fp);
	}
	if (ev) cudaEventRecord(ev[1], stream);
	if (gateWord)
	{
		// strip mode: the tile kernel's stores go to another rank's framebuffer, which must be free first
		k_wait<<<1, 32, 0, stream>>>(gateWord, 1, gateValue);
		pdl = false;
	}
	if (fp.tileRows > 0)
	{
		// k_raster may become resident while k_geom drains (unless stage events sit between)
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.numAttrs = (ev || !pdl) ? 0 : 1;
		cfg.gridDim = dim3(fp.tilesX, fp.tileRows);
		cfg.blockDim = dim3(MR_RASTER_THREADS);
		cfg.dynamicSmemBytes = 0;
		cudaLaunchKernelEx(&cfg, inl ? k_raster<TM_INLINE> : k_raster<TM_GLOBAL>, fp);
	}
	else
	{
		// k_raster normally resets the work-list counters and the next frame's statistics
		cudaMemsetAsync(fp.geomSync, 0, 3 * MR_SYNC_STRIDE * sizeof(int), stream);
		cudaMemsetAsync(fp.ctrNext, 0, sizeof(Counters), stream);
	}
	if (ev) cudaEventRecord(ev[2], stream);
	if (bracketStop) cudaEventRecord(bracketStop, stream);
}

#ifdef MR_TIMELINE
extern "C" __attribute__((visibility("default"))) int mr_debug_raster_timeline(unsigned long long* out, int nWords)
{
	const size_t n = std::min((size_t)nWords, sizeof(g_rtl) / 8);
	return cudaMemcpyFromSymbol(out, g_rtl, n * 8) == cudaSuccess ? (int)n : -1;
}
extern "C" __attribute__((visibility("default"))) int mr_debug_tri_time(unsigned long long* out, int nWords)
{
	const size_t n = std::min((size_t)nWords, sizeof(g_tlTri) / 8);
	return cudaMemcpyFromSymbol(out, g_tlTri, n * 8) == cudaSuccess ? (int)n : -1;
}
extern "C" __attribute__((visibility("default"))) int mr_debug_timeline(unsigned long long* out, int nWords)
{
	const size_t n = std::min((size_t)nWords, sizeof(g_timeline) / 8);
	return cudaMemcpyFromSymbol(out, g_timeline, n * 8) == cudaSuccess ? (int)n : -1;
}
#endif

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

void mrk_launch_clear_rows(float* image, float* depth, int w, int rowBegin, int rowEnd, float r, float g, float b, cudaStream_t stream)
{
	if (rowEnd > rowBegin)
		k_clear_rows<<<148 * 4, 256, 0, stream>>>(image, depth, (size_t)rowBegin * w, (size_t)rowEnd * w, r, g, b);
}

void mrk_launch_signal(unsigned* word, unsigned value, cudaStream_t stream) { k_signal<<<1, 1, 0, stream>>>(word, value); }
void mrk_launch_wait(const unsigned* words, int n, unsigned value, cudaStream_t stream) { k_wait<<<1, 32, 0, stream>>>(words, n, value); }

void mrk_launch_flush_read(const void* buf, size_t bytes, float* sink, cudaStream_t stream)
{
	k_flush_read<<<148 * 8, 256, 0, stream>>>((const float4*)buf, bytes / 16, sink);
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
