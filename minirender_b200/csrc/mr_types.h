// Device-side data model of the B200 rasterization pipeline (shared by kernels and host code).
//
// Scene-static arrays (uploaded by mr_upload_scene, resident in HBM):
//   meshlets[]    one blob per MR_CLUSTER consecutive triangles of a mesh: the distinct (position, normal,
//                 texcoord) corners of those triangles as two float4 planes, then one packed word of three
//                 10-bit local indices per triangle. A blob is what one team of k_geom pulls into shared
//                 memory with a single cp.async.bulk; meshletDir[] gives offset and vertex count
//   clusters[]    bounding sphere + normal cone of the same triangles (cluster culling)
//   texels[]      float4 per texel (r,g,b,0), all textures concatenated
// Per-frame arrays:
//   RStat[]/RDyn[] per renderable (flattened scene entry): mesh + instance bases / matrices
//   MatDev[]       materials
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

struct MeshDev // host-side bookkeeping of one uploaded mesh
{
	int clusterBase; // its first cluster in clusters[] / meshletDir[]
	int nTri;
	int hasUV;
	int pad;
};

#define MR_CLUSTER 32 // triangles per cluster = per meshlet = what one warp of k_geom sets up at a time

// Meshlet blob of one cluster (scene-static, built by mr_upload_scene): nv distinct corners
//   plane 0: nv x float4 (px, py, pz, nx)      object-space position, normal x
//   plane 1: nv x float4 (ny, nz, u, v)        normal y z, texcoord (0,0 when the mesh has none)
//   index  : MR_CLUSTER x uint32               i0 | i1 << 10 | i2 << 20 (local corner indices; padding triangles: 0)
// Blobs start at multiples of 16 bytes and are a multiple of 16 bytes long (cp.async.bulk).
#define MR_MESHLET_MAX_VERTS (3 * MR_CLUSTER)
#define MR_MESHLET_BYTES(nv) ((nv) * 32 + MR_CLUSTER * 4)
#define MR_MESHLET_NONFINITE 1u // a corner position of the meshlet is NaN or infinite
struct MeshletDir
{
	uint32_t off16; // blob offset into meshlets[] in units of 16 bytes
	uint32_t nv;    // corners | MR_MESHLET_* flags << 16
};

struct __align__(16) RStat // per renderable, changes only when the flattened structure changes
{
	int triBase;     // first triangle instance (submission order); a multiple of MR_CLUSTER: every
	                 // cluster lies inside one renderable and maps onto one meshlet of its mesh
	int clusterBase; // the mesh's first cluster in clusters[] / meshletDir[]
	int nTri;        // triangles of the mesh (instances beyond it are padding)
	int triBaseReal; // first triangle instance counted without padding (the reference's submission index)
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

// One unit of work of k_geom: a cluster that survived culling, with everything a warp needs to process it
// (so that taking it is a single 128-byte read). Written by k_geom's own cull phase.
struct __align__(128) GeomEntry
{
	int ci;          // cluster: triangle instances [ci * MR_CLUSTER, (ci + 1) * MR_CLUSTER); -1 = no more work
	int nv;          // corners in its meshlet
	uint32_t off16;  // meshlet blob offset (16-byte units)
	int triFirst;    // its first triangle within the mesh
	int nTri;        // triangles of the mesh
	int subBase;     // RStat::triBaseReal
	int material;
	uint32_t flags;  // MR_MESHLET_*
	float mv[12];
	float nm[12];
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
#define MR_REC_CLIPPED 1u // produced by the near-plane clipper (bits 1..31 of a record's flags word: 1 + its first checkpoint, 0 = none)
#ifndef MR_EXACT_SHADING
#define MR_EXACT_SHADING 0 // 1: shading in the reference's operation order, double pow, corner positions interpolated (float RGB bit-identical)
#endif
#define MR_UNPROJECT_POSITIONS (!MR_EXACT_SHADING)
#define MR_CHK_MIN_TILES 16 // triangles spanning at least this many tile-column boundaries get edge-chain checkpoints
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

#define MR_SYNC_STRIDE 256 // ints between k_geom's global counters
#define MR_STAT_SLOTS 32 // statistics are spread over this many slots (summed by the host): no single-address hot spot
struct Counters
{
	// line 0: written by k_geom, only read by k_raster
	unsigned long long trianglesIn;
	unsigned long long ovfTotal;    // entries appended to the overflow list (may exceed its capacity)
	unsigned int overflow;          // the overflow list did not fit: the frame must be re-run with more room
	unsigned int pad0;
	unsigned int visible, clusters; // k_geom: clusters that survived culling / clusters of the frame
	// edge-chain checkpoints of wide triangles (see k_chain): pool entries / work items handed out this frame, and what
	// the frame would have needed (the host sizes the buffers of the following frames from these)
	unsigned int chkUsed, chkItems;
	unsigned long long chkDemand;
	unsigned int chkItemDemand, pad1a;
	// tiles this frame may have drawn into, as a rectangle, all four as maxima so that 0 = nothing: tilesX - tx0, tx1 + 1,
	// tilesY - ty0, ty1 + 1 (k_geom: the bboxes of everything that emitted a fragment or was binned). Everything outside
	// holds the clear values: a host mirror of the image only needs this rectangle (mr_read_image_dirty_begin).
	unsigned int dirty[4];
	unsigned long long pad1[7];
	// line 1 (offset 128): written by k_raster
	unsigned int maxTile;           // largest per-tile count among tiles that spilled
	unsigned int pad2;
	unsigned long long pad3[15];
	// statistics, one warp-aggregated RED per warp into slot (warp index % MR_STAT_SLOTS)
	unsigned long long records[MR_STAT_SLOTS];   // set-up triangles that can own a pixel
	unsigned long long clippedIn[MR_STAT_SLOTS]; // input triangles crossing the near plane
	unsigned long long zeroCov[MR_STAT_SLOTS];   // set-up small triangles that cover no pixel centre (dropped)
	unsigned long long pairTotal[MR_STAT_SLOTS]; // (tile, triangle) pairs of the frame (summed by the tile kernel)
	unsigned long long tilesStored[MR_STAT_SLOTS]; // tiles the tile kernel wrote to the framebuffer
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
	int tightScan; // small triangles skip the outermost columns / rows of the reference's loops where those provably cover nothing
	int sparseStores; // tiles nothing was drawn into are not written (the target already holds the clear values)
	// Edge-chain checkpoints (k_chain): for a wide triangle, the accumulated edge functions (e1, e2) of every row at
	// every 16-column tile boundary of its bbox, so that a tile starts its replay of the reference's chain
	// (Renderer.cpp:241-243) at its own left edge instead of at the triangle's.
	float2* chkPool;   // [chkCap] entries; a triangle's block is rows x (tile boundaries inside its bbox), row-major
	int4* chkItems;    // [chkItemCap] work items of k_chain: (record id, first pool entry, block of 32 rows, -)
	int chkCap, chkItemCap;
	int chkEnable;     // this frame runs k_chain: k_geom may hand out checkpoints
	int chkMinTiles;   // ... to triangles whose bbox spans at least this many tile-column boundaries
	int stdProj; // standard perspective matrix with the near plane in front of the eye: projectStd() applies
	// Standard perspective form (whatever the near plane): the view-space position of a covered pixel is its own ray
	// scaled by the depth, position = z * (unprojX.x * (j + 0.5) + unprojX.y, unprojY.x * (i + 0.5) + unprojY.y, -1) -
	// the point the reference interpolates from the corners (Renderer.cpp:281), up to rounding, which only shading sees.
	// Records then carry no corner positions (three 32-byte pairs per untextured triangle instead of four).
	int unproject;
	float unprojX[2], unprojY[2];
	int nRenderables, nTriInst;
	int debug; // mr_set_debug flags
	int binCap; // entries per tile bin
	int ovfCap; // entries in the overflow list

	const unsigned char* meshlets; // meshlet blobs (see MeshletDir)
	const MeshletDir* meshletDir;  // per mesh cluster
	const float4* texels;
	const RStat* rstat;
	const RDyn* rdyn;
	const MatDev* mats;
	// Cluster culling: per mesh cluster of MR_CLUSTER triangles, two float4 in object space:
	// (bounding-sphere centre, radius), (normal-cone axis, sin(cone half-angle + margin) or 2 = no cone)
	const float4* clusters;
	const int* triBlockCl; // renderable of every cluster of the frame (MR_CLUSTER triangle instances)
	int nTriReal;          // triangles submitted (nTriInst counts the per-renderable padding too)
	int cullClusters;      // 0: off (orthographic or non-standard projection): every cluster is processed
	float cullPlanes[4][4]; // view-space planes (unit normal, offset) bounding the rows / columns this frame can touch

	// k_geom: persistent CTAs; a warp's shared memory holds one meshlet (geomVertCap corners) and its transformed corners
	int geomVertCap;
	GeomEntry* visEntries; // work list of the frame: clusters that survived culling (any order)
	int* geomSync;         // three counters MR_SYNC_STRIDE ints apart (own L2 lines): [0] chunks of the list handed out,
	                       // [1] entries appended, [2] CTAs that finished culling; zeroed by k_raster

	unsigned long long* gkeys; // per pixel: orderable z << 32 | record index + 1 (MR_KEY_EMPTY: untouched)
	float4* recs;        // records of sub-triangle 0 in plane layout: block (t >> 5), field pair j, lane (t & 31), 32 bytes each
	float4* recs1;       // records of sub-triangle 1 (second clipper output): MR_REC_FIELDS float4 per triangle
	int2* tileCount;     // per tile: x = triangles binned (may exceed binCap: the rest is in ovfPairs),
	                     // y = nonzero if fragments of small triangles may have reached the tile's gkeys.
	                     // All zero between frames: the tile kernel resets what it reads
	int* bins;           // tilesX*tilesY bins of binCap record indices
	int2* ovfPairs;      // (tile, record) entries that did not fit their bin
	Counters* ctr;       // this frame's counters (zero when the frame starts)
	Counters* ctrNext;   // the next frame's: zeroed by k_raster

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
int mrk_geom_config(int nvCap, int smCount, int* grid, int* smemBytes);
void mrk_launch_frame(const FrameParams& fp, int geomGrid, int geomSmem, cudaStream_t stream, cudaEvent_t* stageEvents /* 3 or NULL */,
                      cudaEvent_t bracketStart, cudaEvent_t bracketStop, bool pdl, const unsigned* gateWord, unsigned gateValue);
void mrk_launch_clear_rows(float* image, float* depth, int w, int rowBegin, int rowEnd, float r, float g, float b, cudaStream_t stream);
int mrk_selftest_no_fma(cudaStream_t stream);
void mrk_launch_signal(unsigned* word, unsigned value, cudaStream_t stream);
void mrk_launch_wait(const unsigned* words, int n, unsigned value, cudaStream_t stream);
void mrk_launch_flush_read(const void* buf, size_t bytes, float* sink, cudaStream_t stream);
void mrk_launch_range(const float* depth, float* xyz, int w, int h, const float* P16, cudaStream_t stream);
void mrk_launch_rgb8(const float* image, uint8_t* out, size_t nFloats, cudaStream_t stream);
void mrk_launch_pack(float4* dst4, const float* src, int n, int comps, float w, cudaStream_t stream);

#endif
