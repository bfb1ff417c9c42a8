// Device-side data model of the B200 rasterization pipeline (shared by kernels and host code).
//
// Scene-static arrays (uploaded by mr_upload_scene, resident in HBM):
//   pos4[]   float4 per mesh vertex   (x,y,z,1)      -- one LDG.128 per vertex, 16 B aligned
//   nrm4[]   float4 per mesh normal   (x,y,z,0)
//   uv2[]    float2 per texcoord
//   idxPos[] idxNrm[] idxUv[]  int per triangle corner (3 per triangle)
//   texels[] float4 per texel (r,g,b,0), all textures concatenated
//   MeshDev[] per mesh: base offsets into the arrays above
// Per-frame arrays:
//   RStat[]/RDyn[] per renderable (flattened scene entry): mesh + instance bases / matrices
//   MatDev[]       materials
//   pv[]           float4 per vertex *instance*: (pixel x, pixel y, view z, depth term)
//   recs[]         one 160-byte record per set-up triangle (10 float4 fields: 4 raster, 6 shading = its
//                  three view-space corners), at index 2*t+sub where t is the triangle instance index in
//                  submission order: the index IS the submission id that resolves equal-depth ties.
//                  Stored as five 32-byte pairs (256-bit accesses), in planes of 32 triangles
//   gkeys[]        one 64-bit depth key per pixel: small triangles depth-test straight into it
//                  (atomicMin in L2); the tile kernel merges, resolves and resets it every frame
//   tileCount[] bins[] ovfPairs[]   16x16-tile binning of the larger triangles: fixed-capacity
//                  bins + global overflow list
#ifndef MR_TYPES_H
#define MR_TYPES_H

#include <cuda_runtime.h>
#include <stdint.h>

#define MR_TILE 16
#define MR_TILE_SHIFT 4
#define MR_TILE_PIXELS 256
#define MR_SEG_PER_LANE 4 // tiles per triangle binned on the warp-aggregated fast path

struct MeshDev
{
	int posBase, nrmBase, uvBase; // into pos4 / nrm4 / uv2
	int triBase;                  // first triangle in idxPos/idxNrm (units: triangles)
	int uvTriBase;                // first triangle in idxUv, or -1 when the mesh has no uv stream
	int nPos, nTri, hasUV;
};

#define MR_CLUSTER 128 // triangles per cull cluster = triangles per k_setup CTA

struct __align__(16) RStat // per renderable, changes only when the flattened structure changes
{
	int vertBase;   // first vertex instance (into pv)
	int triBase;    // first triangle instance (submission order); a multiple of MR_CLUSTER: every
	                // k_setup CTA lies inside one renderable and maps onto one cluster of its mesh
	int nrmBase;    // first normal instance (into vnrm4)
	int idxBase;    // the mesh's first triangle in idxPos / idxNrm
	int posBase;    // the mesh's first vertex in pos4
	int nrmSrcBase; // the mesh's first normal in nrm4
	int uvBase;     // the mesh's first texcoord in uv2
	int uvTriBase;  // the mesh's first triangle in idxUv, -1: no texcoords
	int clusterBase; // the mesh's first cluster in clusters[]
	int nTri;        // triangles of the mesh (instances beyond it are padding)
	int triBaseReal; // first triangle instance counted without padding (the reference's submission index)
	int pad;
};

struct __align__(16) RDyn // per renderable, per frame
{
	float mv[12]; // modelview, row-major 3x4
	float nm[12]; // normal matrix, row-major 3x4
	int material;
	int cullFlags;     // bit 0: the modelview is a similarity with positive determinant (normal cones stay cones)
	float radiusScale; // upper bound of the modelview's stretch: scales a cluster's bounding radius
	int pad;
};

struct __align__(16) MatDev
{
	float diffuse[3];
	float shininess;
	float specular[3];
	int texOffset; // into texels, -1 = no texture
	float emissive[3];
	int texRows;
	int texCols;
	int pad[3];
};

// 64-byte raster record: everything coverage + depth need (reference Renderer.cpp:212-224).
#define MR_REC_FIELDS 10 // float4 fields per record: 4 raster (struct Rec) + 6 shading (struct ShadeRec)
#define MR_REC_CLIPPED 1u // produced by the near-plane clipper
#define MR_KEY_EMPTY 0xffffffffffffffffull // gkeys[] entry no fragment has touched
struct __align__(16) Rec
{
	float p0x, p0y, p2x, p2y; // edge origins (e1 is measured from p2, e2 from p0)
	float n1x, n1y, n2x, n2y; // scaled edge normals
	float d0, d1, d2;         // iz[] (perspective) or zz[] (ortho)
	uint32_t material;        // into mats
	uint32_t xspan;           // x0 | x1 << 16 : first / last pixel column of the clamped bbox
	uint32_t yspan;           // y0 | y1 << 16
	uint32_t flags;
	uint32_t submission; // 2 * unpadded triangle instance + clipper output: the id mr_read_winner_ids reports
};

// 96-byte shading record of the same triangle: the three corners in view space (for clipper output:
// the clipped corners), i.e. what the reference's paintTriangle receives (Renderer.cpp:351-380).
// Shading a pixel reaches everything it needs in one hop from the winner's id.
struct __align__(16) ShadeRec
{
	float p0[3], u0; // view-space position of corner 0, texcoord u of corner 0
	float p1[3], v0;
	float p2[3], u1;
	float n0[3], v1; // view-space normal of corner 0
	float n1[3], u2;
	float n2[3], v2;
};

#define MR_STAT_SLOTS 32 // statistics are spread over this many slots (summed by the host): no single-address hot spot
struct Counters
{
	// line 0: written by k_vertex / k_setup, only read by k_raster
	unsigned long long trianglesIn;
	unsigned long long ovfTotal;    // entries appended to the overflow list (may exceed its capacity)
	unsigned int overflow;          // the overflow list did not fit: the frame must be re-run with more room
	unsigned int pad0;
	unsigned int visible, clusters; // k_setup: clusters that survived culling / clusters of the frame (sizes the next frame's grid)
	unsigned long long pad1[12];
	// line 1 (offset 128): written by k_raster
	unsigned int maxTile;           // largest per-tile count among tiles that spilled
	unsigned int pad2;
	unsigned long long pad3[15];
	// statistics, one warp-aggregated RED per warp into slot (warp index % MR_STAT_SLOTS)
	unsigned long long records[MR_STAT_SLOTS];   // set-up triangles that can own a pixel
	unsigned long long clippedIn[MR_STAT_SLOTS]; // input triangles crossing the near plane
	unsigned long long zeroCov[MR_STAT_SLOTS];   // set-up small triangles that cover no pixel centre (dropped)
	unsigned long long pairTotal[MR_STAT_SLOTS]; // (tile, triangle) pairs of the frame (summed by the tile kernel)
};

#define MR_INLINE_TABLE 32 // renderables / materials that travel inside the kernel parameters

struct FrameParams
{
	float P[16];
	float light[3];
	float ambient;
	float bg[3];
	float znear;
	float wf, hf;
	int w, h;
	int tilesX, tilesY;
	int tileRow0, tileRows; // tile rows covered by this frame (strip rendering)
	// Tiles tx < fullTx, fullTy0 <= ty < fullTy1 lie fully inside the image and the strip and may be stored as
	// whole 32-byte sectors (width a multiple of 8, not an accumulating frame): no per-pixel bounds tests there.
	int fullTx, fullTy0, fullTy1;
	float bgPattern[12];    // r g b r g b ...
	int rowBegin, rowEnd;   // pixel rows [rowBegin,rowEnd)
	int persp, lightIsPoint, lighting, texturing, saveNormals, keep;
	int nRenderables, nVertInst, nTriInst;
	int debug; // mr_set_debug flags
	int binCap; // entries per tile bin
	int ovfCap; // entries in the overflow list

	const float4* pos4;
	const float4* nrm4;
	const float2* uv2;
	const int* idxPos;
	const int* idxNrm;
	const int* idxUv;
	const float4* texels;
	const MeshDev* meshes;
	const RStat* rstat;
	const RDyn* rdyn;
	const MatDev* mats;
	// Cluster culling (k_setup): per mesh cluster of MR_CLUSTER triangles, two float4 in object space:
	// (bounding-sphere centre, radius), (normal-cone axis, sin(cone half-angle + margin) or 2 = no cone)
	const float4* clusters;
	const int* triBlockCl; // renderable of every k_setup CTA (MR_CLUSTER triangle instances)
	unsigned* clusterVis;  // a bit per cluster: 0 = culled this frame (written by k_vertex, one word per warp)
	int* visList;          // clusters that can reach a pixel of this frame (written by k_vertex, any order)
	int* visCount;         // their number; zeroed again by k_raster once k_setup has consumed the list
	int setupCtas;         // persistent k_setup CTAs
	int nNrmSrc;           // float4 elements of nrm4 that may be prefetched
	int nTriReal;          // triangles submitted (nTriInst counts the per-renderable padding too)
	int cullClusters;      // 0: off (orthographic or non-standard projection)
	float cullPlanes[4][4]; // view-space planes (unit normal, offset) bounding the rows / columns this frame can touch
	const int* vtxBlockR; // renderable that owns the first vertex instance of each 256-block

	float4* pv;          // per vertex instance: pixel x, pixel y, view z, depth term
	unsigned long long* gkeys; // per pixel: orderable z << 32 | record index + 1 (MR_KEY_EMPTY: untouched)
	float4* recs;        // records of sub-triangle 0 in plane layout: block (t >> 5), field pair j, lane (t & 31), 32 bytes each
	float4* recs1;       // records of sub-triangle 1 (second clipper output): MR_REC_FIELDS float4 per triangle
	int2* tileCount;     // per tile: x = triangles binned (may exceed binCap: the rest is in ovfPairs),
	                     // y = nonzero if fragments of small triangles may have reached the tile's gkeys
	int* bins;           // tilesX*tilesY bins of binCap record indices
	int2* ovfPairs;      // (tile, record) entries that did not fit their bin
	Counters* ctr;

	// Small scenes: the per-frame tables ride in the kernel parameters (no H2D copy per frame).
	// rdyn / mats point at these arrays then (set up on the device: see frameTables()).
	int inlineTables;
	RStat rstatInline[MR_INLINE_TABLE];
	RDyn rdynInline[MR_INLINE_TABLE];
	MatDev matsInline[MR_INLINE_TABLE];

	float* image;   // h*w*3
	float* depth;   // h*w
	float* normals; // h*w*3 or NULL
	int* winner;    // h*w submission ids (debug) or NULL
};

// kernel launchers (mr_kernels.cu)
void mrk_launch_frame(const FrameParams& fp, cudaStream_t stream, cudaEvent_t* stageEvents /* 6 or NULL */, cudaEvent_t bracketStart = 0,
                      cudaEvent_t bracketStop = 0);
int mrk_selftest_no_fma(cudaStream_t stream);
void mrk_launch_flush_read(const void* buf, size_t bytes, float* sink, cudaStream_t stream);
void mrk_launch_range(const float* depth, float* xyz, int w, int h, const float* P16, cudaStream_t stream);
void mrk_launch_rgb8(const float* image, uint8_t* out, size_t nFloats, cudaStream_t stream);
void mrk_launch_pack(float4* dst4, const float* src, int n, int comps, float w, cudaStream_t stream);

#endif
