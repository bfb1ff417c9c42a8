// C ABI of the B200 rasterization path (include/minirender_b200.h): context, device memory,
// scene upload, per-frame tables and kernel launches. Host code only; kernels are in
// mr_kernels.cu. There is deliberately no CPU path here: every entry point that produces pixels
// needs a CUDA device and fails with MR_E_NO_DEVICE / MR_E_CUDA otherwise.
#include "../../include/minirender_b200.h"
#include "mr_types.h"
#include "../host/HostPool.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct DevBuf
{
	void* p;
	size_t cap;
	DevBuf() : p(0), cap(0) {}
	cudaError_t ensure(size_t bytes, bool exact = false)
	{
		if (bytes <= cap)
			return cudaSuccess;
		if (p)
			cudaFree(p);
		p = 0;
		cap = 0;
		const size_t want = exact ? bytes : bytes + bytes / 4 + 256;
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess)
			cap = want;
		return e;
	}
	void release()
	{
		if (p)
			cudaFree(p);
		p = 0;
		cap = 0;
	}
	template <class T> T* as() const { return (T*)p; }
};

struct PinBuf
{
	void* p;
	size_t cap;
	PinBuf() : p(0), cap(0) {}
	cudaError_t ensure(size_t bytes)
	{
		if (bytes <= cap)
			return cudaSuccess;
		if (p)
			cudaFreeHost(p);
		p = 0;
		cap = 0;
		const size_t want = bytes + bytes / 2 + 4096;
		cudaError_t e = cudaMallocHost(&p, want);
		if (e == cudaSuccess)
			cap = want;
		return e;
	}
	void release()
	{
		if (p)
			cudaFreeHost(p);
		p = 0;
		cap = 0;
	}
};

}

struct mr_ctx
{
	int device;
	cudaStream_t stream;
	bool ownStream;
	cudaStream_t aux; // counter read-back, off the render stream
	std::string error;

	int w, h, tilesX, tilesY;
	int smCount;

	// scene-static device arrays
	DevBuf meshlets, meshletDir, texels, clusters, triBlockCl, visEntries, geomSync;
	int geomVertCap, geomGrid, geomSmem; // shape of k_geom for this scene's meshlets
	std::vector<int> hostClusterBase; // first cluster of every mesh
	std::vector<MeshDev> hostMeshes;
	std::vector<int> texOffset, texRows, texCols;
	bool haveScene;
	unsigned sceneSerial;

	// per-frame tables
	DevBuf rstat, rdyn, mats;
	std::vector<int> structureKey; // mesh id per renderable of the tables currently on the device
	unsigned structureSerial;
	int nTriInst;
	// per-frame host staging + counters, a ring so that mr_render never waits for the GPU
	struct Slot
	{
		PinBuf stage;
		Counters* hostCtr; // pinned
		cudaEvent_t kernelsDone; // on the render stream, after the frame's last kernel
		cudaEvent_t done;        // on the aux stream, after the counters reached the host
		bool pending;
		int dirty[4];      // pixel rectangle [x0, x1) x [y0, y1) the frame may have drawn into (from its counters; empty: x1 <= x0)
		bool wholeFrame;   // the frame wrote every pixel of the image with clear values outside `dirty` (no keep, no row range)
		float bg[3];
		Slot() : hostCtr(0), kernelsDone(0), done(0), pending(false), wholeFrame(false) { dirty[0] = dirty[1] = dirty[2] = dirty[3] = 0; bg[0] = bg[1] = bg[2] = 0; }
	};
	enum { kSlots = 4 };
	Slot slots[kSlots];
	int slotNext;   // slot the next frame uses
	int slotNewest; // slot of the most recent frame (-1: none)

	// scratch
	DevBuf recs, recs1, tileCount, ovfPairs, bins, ctr, gkeys;
	DevBuf chkPool, chkItems; // edge-chain checkpoints of wide triangles (k_chain): sized from the previous frames' demand
	size_t chkWantEntries, chkWantItems;
	bool chkActive;           // the most recently retired frame had wide triangles: the following frames run k_chain
	bool slotOverflowed; // a frame older than the newest one overflowed its spill list (async readers are told)
	bool noClusterCull, noPdl, noStdProj, noTightScan, noChain; // MR_NO_CLUSTER_CULL / MR_NO_PDL in the environment when the context was created
	int binCap;    // entries per tile bin
	int binCapWanted;
	size_t ovfCap; // entries in the overflow list
	size_t h2dBytesLastFrame;

	// outputs
	// outputs: one or two sets of image/depth buffers (mr_set_output_slots); `image`/`depth` below always
	// denote the set the newest frame went to
	DevBuf imageSlot[2], depthSlot[2], normals, winner, scratchOut, flushBuf, syncWords;
	int outSlots, outCur;
	cudaStream_t copy;          // device->host copies that overlap the next frame (mr_read_image_begin)
	cudaEvent_t frameDone[2];   // render stream: the frame in slot s is complete
	cudaEvent_t copyDone[2];    // copy stream: the host copy out of slot s is complete
	bool copyPending[2];
	int copyRing[2];            // frame-ring slot of the frame whose image the pending copy of output set s reads
	// host buffers filled by mr_read_image_dirty_begin: what rectangle of each is not background, so that the next
	// frame copied into it only needs the union of the old and the new rectangle
	struct HostMirror { const void* ptr; int w, h; float bg[3]; int rect[4]; };
	std::vector<HostMirror> mirrors;
	void *remoteImage, *remoteDepth;
	bool sparseRemote;          // mr_set_sparse_remote_stores
	const unsigned* gateWord;   // mr_set_raster_gate (one shot)
	unsigned gateValue;
	int debugFlags;
	cudaEvent_t timingStart, timingStop; // mr_set_timing_events: recorded around the next frame's launches

	// last frame (kept for overflow re-runs and profiling)
	mr_frame lastFrame;
	std::vector<mr_renderable> lastRenderables;
	std::vector<mr_material> lastMaterials;
	bool haveFrame;
	mr_stats stats;

	mr_ctx() : device(0), stream(0), ownStream(false), aux(0), w(0), h(0), tilesX(0), tilesY(0), haveScene(false), sceneSerial(0),
	           structureSerial(~0u), geomVertCap(0), geomGrid(0), geomSmem(0), nTriInst(0), slotOverflowed(false), noClusterCull(false), noPdl(false), noStdProj(false), noTightScan(false), noChain(false), chkWantEntries(0), chkWantItems(0), chkActive(false), slotNext(0), slotNewest(-1), binCap(0), binCapWanted(0), ovfCap(0), h2dBytesLastFrame(0), remoteImage(0), sparseRemote(false), gateWord(0), gateValue(0),
	           remoteDepth(0), debugFlags(0), timingStart(0), timingStop(0), haveFrame(false), outSlots(1), outCur(0), copy(0)
	{
		frameDone[0] = frameDone[1] = copyDone[0] = copyDone[1] = 0;
		copyPending[0] = copyPending[1] = false;
		copyRing[0] = copyRing[1] = -1;
		memset(&lastFrame, 0, sizeof(lastFrame));
		memset(&stats, 0, sizeof(stats));
	}
};

namespace {

int setError(mr_ctx* c, int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (c)
		c->error = buf;
	return code;
}

#define MR_CUDA(c, call)                                                                                       \
	do                                                                                                         \
	{                                                                                                          \
		cudaError_t e_ = (call);                                                                               \
		if (e_ != cudaSuccess)                                                                                 \
			return setError(c, e_ == cudaErrorMemoryAllocation ? MR_E_NOMEM : MR_E_CUDA, "%s: %s", #call,      \
			                cudaGetErrorString(e_));                                                           \
	} while (0)

struct Bind
{
	int prev;
	bool changed;
	explicit Bind(int dev) : prev(-1), changed(false)
	{
		if (cudaGetDevice(&prev) == cudaSuccess && prev != dev)
		{
			cudaSetDevice(dev);
			changed = true;
		}
	}
	~Bind()
	{
		if (changed)
			cudaSetDevice(prev);
	}
};

int launchFrame(mr_ctx* c, const mr_frame* f, cudaEvent_t* ev, bool rerun = false);

void absorbCounters(mr_ctx* c, const Counters& k)
{
	c->stats.triangles_in = (int64_t)k.trianglesIn;
	unsigned long long rec = 0, clip = 0, pairs = 0, zero = 0, stored = 0;
	for (int i = 0; i < MR_STAT_SLOTS; i++)
	{
		rec += k.records[i];
		clip += k.clippedIn[i];
		pairs += k.pairTotal[i];
		stored += k.tilesStored[i];
		zero += k.zeroCov[i];
	}
	c->stats.records = (int64_t)rec;
	c->stats.clipped_in = (int64_t)clip;
	c->stats.bin_entries = (int64_t)pairs;
	c->stats.zero_coverage = (int64_t)zero;
	c->stats.tiles_stored = (int64_t)stored;
	c->stats.clusters = (int64_t)k.clusters;
	c->stats.clusters_visible = (int64_t)k.visible;
	c->stats.tiles_x = c->tilesX;
	c->stats.tiles_y = c->tilesY;
}

// Retires one slot: waits for its frame, reads its counters. Returns 1 if that frame overflowed
// its overflow list (then the capacities have been raised), 0 if fine, <0 on error.
int retireSlot(mr_ctx* c, int i)
{
	mr_ctx::Slot& s = c->slots[i];
	if (!s.pending)
		return 0;
	MR_CUDA(c, cudaEventSynchronize(s.done));
	s.pending = false;
	const Counters k = *s.hostCtr;
	absorbCounters(c, k);
	{
		const int T = MR_TILE;
		s.dirty[0] = k.dirty[0] ? std::max(0, (c->tilesX - (int)k.dirty[0]) * T) : 0;
		s.dirty[1] = std::min(c->w, (int)k.dirty[1] * T);
		s.dirty[2] = k.dirty[2] ? std::max(0, (c->tilesY - (int)k.dirty[2]) * T) : 0;
		s.dirty[3] = std::min(c->h, (int)k.dirty[3] * T);
	}
	// Many spilled entries make the tile kernel scan a long overflow list: give the bins more room
	// for the following frames (the current frame is still correct).
	if (k.ovfTotal > 4096 && k.maxTile > (unsigned)c->binCap)
	{
		int want = c->binCap;
		while (want < (int)std::min<unsigned>(k.maxTile, 1u << 20))
			want <<= 1;
		c->binCapWanted = std::max(c->binCapWanted, want);
	}
	// wide triangles: their checkpoint demand decides whether the next frames run k_chain and with how much room
	// (k_chain costs a launch and one serial walk over the widest row, ~4-8 us: it pays in frames dominated by triangles
	// hundreds of pixels wide - a floor, walls, a backdrop - measured 312 -> 139 us of k_raster on two dozen of them)
	c->chkActive = k.chkDemand >= (1u << 16);
	if (k.chkDemand > 0)
	{
		c->chkWantEntries = std::max(c->chkWantEntries, (size_t)std::min<unsigned long long>(k.chkDemand + k.chkDemand / 4 + 4096, 0x7ff00000ull));
		c->chkWantItems = std::max(c->chkWantItems, (size_t)k.chkItemDemand + (size_t)k.chkItemDemand / 4 + 1024);
	}
	c->stats.chk_entries = (int64_t)std::min<unsigned long long>(k.chkUsed, c->chkPool.cap / sizeof(float2));
	c->stats.chk_demand = (int64_t)k.chkDemand;
	if (!k.overflow)
		return 0;
	c->ovfCap = std::max(c->ovfCap, (size_t)k.ovfTotal + (size_t)k.ovfTotal / 4 + 1024);
	c->stats.regrows++;
	return 1;
}

// Waits for everything in flight. If the most recent frame overflowed its pair queues it is
// rendered again with more room (earlier frames have been overwritten by then anyway).
int finishFrame(mr_ctx* c)
{
	for (int attempt = 0; attempt < 5; attempt++)
	{
		int newestOverflowed = 0;
		for (int n = 0; n < mr_ctx::kSlots; n++)
		{
			const int i = (c->slotNext + n) % mr_ctx::kSlots; // oldest first
			const int rc = retireSlot(c, i);
			if (rc < 0)
				return rc;
			if (i == c->slotNewest)
				newestOverflowed = rc;
		}
		if (!newestOverflowed || !c->haveFrame)
		{
			MR_CUDA(c, cudaStreamSynchronize(c->stream));
			return MR_OK;
		}
		mr_frame f = c->lastFrame;
		f.renderables = c->lastRenderables.empty() ? 0 : &c->lastRenderables[0];
		f.materials = c->lastMaterials.empty() ? 0 : &c->lastMaterials[0];
		// the aborted frame left fragments in the depth keys
		MR_CUDA(c, cudaMemsetAsync(c->gkeys.p, 0xff, (size_t)c->w * c->h * 8, c->stream));
		int rc = launchFrame(c, &f, 0, true);
		if (rc)
			return rc;
	}
	return setError(c, MR_E_OVERFLOW, "pair queue kept overflowing");
}

int ensureOutputs(mr_ctx* c, bool normals, bool winner)
{
	const size_t npix = (size_t)c->w * c->h;
	for (int sl = 0; sl < c->outSlots; sl++)
	{
		const bool freshImage = c->imageSlot[sl].cap < npix * 12;
		MR_CUDA(c, c->imageSlot[sl].ensure(npix * 12, true));
		MR_CUDA(c, c->depthSlot[sl].ensure(npix * 4, true));
		if (freshImage)
		{
			MR_CUDA(c, cudaMemsetAsync(c->imageSlot[sl].p, 0, npix * 12, c->stream));
			MR_CUDA(c, cudaMemsetAsync(c->depthSlot[sl].p, 0, npix * 4, c->stream));
		}
	}
	// per-pixel depth keys: all MR_KEY_EMPTY between frames (the tile kernel resets what it reads)
	if (c->gkeys.cap < npix * 8)
	{
		MR_CUDA(c, c->gkeys.ensure(npix * 8, true));
		MR_CUDA(c, cudaMemsetAsync(c->gkeys.p, 0xff, npix * 8, c->stream));
	}
	if (normals && c->normals.cap < npix * 12)
	{
		MR_CUDA(c, c->normals.ensure(npix * 12, true));
		MR_CUDA(c, cudaMemsetAsync(c->normals.p, 0, npix * 12, c->stream));
	}
	if (winner && c->winner.cap < npix * 4)
	{
		MR_CUDA(c, c->winner.ensure(npix * 4, true));
		MR_CUDA(c, cudaMemsetAsync(c->winner.p, 0xff, npix * 4, c->stream));
	}
	return MR_OK;
}

// Frames with at least kHostPoolMin renderables deal their per-renderable host loops to the library's host threads
// (host/HostPool.h) in pieces of kHostChunk.
enum { kHostPoolMin = 1024, kHostChunk = 256 };

// A renderable's per-frame entry of the device table: matrices, material, and what the cluster cull needs to know
// about the modelview.
static void fillDynamic(RDyn& d, const mr_renderable& r)
{
	memcpy(d.mv, r.modelview, sizeof(float) * 12);
	memcpy(d.nm, r.normalmat, sizeof(float) * 12);
	d.material = r.material;
	{
		// Is the modelview a similarity (uniform scale x rotation) with positive determinant? Then a
		// cluster's normal cone is still a cone in view space and its bounding radius scales by s.
		const float* m = d.mv;
		const double c0[3] = { m[0], m[4], m[8] }, c1[3] = { m[1], m[5], m[9] }, c2[3] = { m[2], m[6], m[10] };
		const double l0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2], l1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2],
		             l2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
		const double d01 = c0[0] * c1[0] + c0[1] * c1[1] + c0[2] * c1[2], d02 = c0[0] * c2[0] + c0[1] * c2[1] + c0[2] * c2[2],
		             d12 = c1[0] * c2[0] + c1[1] * c2[1] + c1[2] * c2[2];
		const double det = c0[0] * (c1[1] * c2[2] - c1[2] * c2[1]) - c0[1] * (c1[0] * c2[2] - c1[2] * c2[0]) + c0[2] * (c1[0] * c2[1] - c1[1] * c2[0]);
		const double lmax = std::max(l0, std::max(l1, l2)), lmin = std::min(l0, std::min(l1, l2));
		const double tol = 1e-4 * lmax;
		const bool similar = lmax > 0 && (lmax - lmin) <= tol && fabs(d01) <= tol && fabs(d02) <= tol && fabs(d12) <= tol && lmax == lmax;
		d.cullFlags = (similar && det > 0) ? 1 : 0;
		// the largest stretch of the linear part: s for a similarity, the Frobenius norm otherwise
		const double stretch = similar ? sqrt(lmax) : sqrt(l0 + l1 + l2);
		d.radiusScale = (stretch == stretch) ? (float)(stretch * 1.0001) : INFINITY;
		d.pad = 0;
	}
}

int launchFrame(mr_ctx* c, const mr_frame* f, cudaEvent_t* ev, bool rerun)
{
	const int nR = f->n_renderables;
	// ---- validate + instance bases ----
	std::vector<RStat> rs((size_t)nR);
	long long tb = 0, triReal = 0;
	bool sameStructure = (c->structureSerial == c->sceneSerial) && ((int)c->structureKey.size() == nR);
	for (int i = 0; i < nR; i++)
	{
		const mr_renderable& r = f->renderables[i];
		if (r.mesh < 0 || r.mesh >= (int)c->hostMeshes.size())
			return setError(c, MR_E_INVALID, "renderable %d: mesh index %d out of range", i, r.mesh);
		if (r.material < 0 || r.material >= f->n_materials)
			return setError(c, MR_E_INVALID, "renderable %d: material index %d out of range", i, r.material);
		const MeshDev& hm = c->hostMeshes[r.mesh];
		rs[i].triBase = (int)tb;
		rs[i].clusterBase = hm.clusterBase;
		rs[i].nTri = hm.nTri;
		rs[i].triBaseReal = (int)triReal;
		tb += ((long long)hm.nTri + MR_CLUSTER - 1) / MR_CLUSTER * MR_CLUSTER; // whole clusters per renderable
		triReal += hm.nTri;
		if (sameStructure && c->structureKey[i] != r.mesh)
			sameStructure = false;
	}
	if (tb > 0x3fffffffLL)
		return setError(c, MR_E_INVALID, "frame too large: %lld triangle instances", tb);
	c->nTriInst = (int)tb;

	// ---- device buffers ----
	const int nTiles = c->tilesX * c->tilesY;
	MR_CUDA(c, c->rstat.ensure(sizeof(RStat) * (size_t)std::max(nR, 1)));
	MR_CUDA(c, c->rdyn.ensure(sizeof(RDyn) * (size_t)std::max(nR, 1)));
	MR_CUDA(c, c->mats.ensure(sizeof(MatDev) * (size_t)std::max(f->n_materials, 1)));
	const int nCB = c->nTriInst / MR_CLUSTER;
	MR_CUDA(c, c->triBlockCl.ensure(sizeof(int) * (size_t)std::max(nCB, 1)));
	MR_CUDA(c, c->visEntries.ensure(sizeof(GeomEntry) * (size_t)std::max(nCB, 1)));
	if (!c->geomSync.p)
	{
		MR_CUDA(c, c->geomSync.ensure(sizeof(int) * 3 * MR_SYNC_STRIDE, true));
		MR_CUDA(c, cudaMemsetAsync(c->geomSync.p, 0, sizeof(int) * 3 * MR_SYNC_STRIDE, c->stream));
	}
	MR_CUDA(c, c->recs.ensure(sizeof(float4) * MR_REC_FIELDS * 32 * (size_t)((c->nTriInst + 31) / 32 + 1)));
	MR_CUDA(c, c->recs1.ensure(sizeof(float4) * MR_REC_FIELDS * (size_t)std::max(c->nTriInst, 1)));
	if (c->tileCount.cap < sizeof(int2) * (size_t)(nTiles + 1))
	{
		// all zero between frames: k_raster resets the entries it reads
		MR_CUDA(c, c->tileCount.ensure(sizeof(int2) * (size_t)(nTiles + 1)));
		MR_CUDA(c, cudaMemsetAsync(c->tileCount.p, 0, c->tileCount.cap, c->stream));
	}
	if (!c->ctr.p)
	{
		// a frame's statistics start from zero: k_raster clears the next frame's slot
		MR_CUDA(c, c->ctr.ensure(sizeof(Counters) * mr_ctx::kSlots, true));
		MR_CUDA(c, cudaMemsetAsync(c->ctr.p, 0, sizeof(Counters) * mr_ctx::kSlots, c->stream));
	}
	{
		// Bin capacity: a power of two, at least 256 and at least 8x the mean triangles per tile,
		// within a 1 GiB budget for the bin array; doubled on demand when tiles spill a lot.
		int want = 256;
		const long long mean8 = 8LL * c->nTriInst / std::max(nTiles, 1);
		while (want < mean8 && want < (1 << 20))
			want <<= 1;
		want = std::max(want, c->binCapWanted);
		while (want > 256 && (long long)want * nTiles * 4 > (1LL << 30))
			want >>= 1;
		c->binCap = std::max(c->binCap, want);
		if (c->ovfCap < 65536)
			c->ovfCap = 65536;
		if (c->ovfCap > 0x7fffffffULL)
			return setError(c, MR_E_OVERFLOW, "more than 2^31 spilled (tile, triangle) pairs");
		MR_CUDA(c, c->bins.ensure(sizeof(int) * (size_t)c->binCap * (size_t)nTiles));
		MR_CUDA(c, c->ovfPairs.ensure(sizeof(int2) * c->ovfCap));
	}
	const bool chkEnable = (c->chkActive || (c->debugFlags & 128)) && !c->noChain && !(c->debugFlags & 256);
	if (chkEnable)
	{
		// at most 1 GiB of checkpoints (a triangle that does not fit walks its chain in the tile kernel, as without them)
		const size_t entries = std::min<size_t>(std::max<size_t>(c->chkWantEntries, (size_t)1 << 20), ((size_t)1 << 30) / sizeof(float2));
		const size_t items = std::min<size_t>(std::max<size_t>(c->chkWantItems, (size_t)1 << 16), (size_t)1 << 24);
		MR_CUDA(c, c->chkPool.ensure(sizeof(float2) * entries));
		MR_CUDA(c, c->chkItems.ensure(sizeof(int4) * items));
	}
	int rc = ensureOutputs(c, f->save_normals != 0, (c->debugFlags & 1) != 0);
	if (rc)
		return rc;
	if (c->outSlots > 1 && !f->keep && !rerun)
		c->outCur = (c->outCur + 1) % c->outSlots; // a cleared frame goes to the other output set
	if (c->copyPending[c->outCur])
	{
		// a host copy still reading the target set must finish before the kernels write into it
		MR_CUDA(c, cudaStreamWaitEvent(c->stream, c->copyDone[c->outCur], 0));
		c->copyPending[c->outCur] = false;
	}

	// ---- stage per-frame tables in pinned memory, one async copy each ----
	const size_t szStat = sameStructure ? 0 : sizeof(RStat) * (size_t)nR;
	const size_t szCB = sameStructure ? 0 : sizeof(int) * (size_t)nCB;
	const size_t szDyn = sizeof(RDyn) * (size_t)nR;
	const size_t szMat = sizeof(MatDev) * (size_t)f->n_materials;
	const size_t total = szStat + szCB + szDyn + szMat + 64;
	const int slotIndex = c->slotNext;
	{
		// this frame's slot and the next one (whose statistics this frame's k_raster clears) must have
		// delivered their counters; both are normally long finished
		const int rc0 = retireSlot(c, slotIndex);
		if (rc0 < 0)
			return rc0;
		if (rc0 > 0)
			c->slotOverflowed = true;
		const int rc1 = retireSlot(c, (slotIndex + 1) % mr_ctx::kSlots);
		if (rc1 < 0)
			return rc1;
		if (rc1 > 0)
			c->slotOverflowed = true;
	}
	mr_ctx::Slot& slot = c->slots[slotIndex];
	MR_CUDA(c, slot.stage.ensure(total));
	char* sp = (char*)slot.stage.p;
	size_t off = 0;
	if (!sameStructure)
	{
		memcpy(sp + off, rs.data(), szStat);
		if (szStat) MR_CUDA(c, cudaMemcpyAsync(c->rstat.p, sp + off, szStat, cudaMemcpyHostToDevice, c->stream));
		off += szStat;
		int* cbr = (int*)(sp + off);
		for (int i = 0; i < nR; i++)
		{
			const int b0 = rs[i].triBase / MR_CLUSTER, b1 = (i + 1 < nR ? rs[i + 1].triBase : c->nTriInst) / MR_CLUSTER;
			for (int b = b0; b < b1; b++)
				cbr[b] = i;
		}
		if (szCB) MR_CUDA(c, cudaMemcpyAsync(c->triBlockCl.p, sp + off, szCB, cudaMemcpyHostToDevice, c->stream));
		off += szCB;
		c->structureKey.resize((size_t)nR);
		for (int i = 0; i < nR; i++)
			c->structureKey[i] = f->renderables[i].mesh;
		c->structureSerial = c->sceneSerial;
	}
	off = (off + 15) & ~(size_t)15;
	RDyn* rd = (RDyn*)(sp + off);
	{
		// (independent per renderable: frames with thousands of them share the loop out, host/HostPool.h)
		struct Dynamic
		{
			RDyn* rd; const mr_renderable* in; int nR;
			void operator()(int chunk) const
			{
				for (int i = chunk * kHostChunk; i < std::min(nR, (chunk + 1) * kHostChunk); i++)
					fillDynamic(rd[i], in[i]);
			}
		} perRenderable = { rd, f->renderables, nR };
		minirender::hostpool::parallelFor(nR >= kHostPoolMin ? minirender::hostpool::Pool::get() : 0, (nR + kHostChunk - 1) / kHostChunk, perRenderable);
	}
	const bool inlineTables = nR <= MR_INLINE_TABLE && f->n_materials <= MR_INLINE_TABLE;
	if (szDyn && !inlineTables) MR_CUDA(c, cudaMemcpyAsync(c->rdyn.p, rd, szDyn, cudaMemcpyHostToDevice, c->stream));
	off += szDyn;
	off = (off + 15) & ~(size_t)15;
	MatDev* md = (MatDev*)(sp + off);
	for (int i = 0; i < f->n_materials; i++)
	{
		const mr_material& m = f->materials[i];
		memcpy(md[i].diffuse, m.diffuse, 12);
		memcpy(md[i].specular, m.specular, 12);
		memcpy(md[i].emissive, m.emissive, 12);
		md[i].shininess = m.shininess;
		md[i].texOffset = -1;
		md[i].texRows = md[i].texCols = 0;
		if (m.texture >= 0)
		{
			if (m.texture >= (int)c->texOffset.size())
				return setError(c, MR_E_INVALID, "material %d: texture index %d out of range", i, m.texture);
			md[i].texOffset = c->texOffset[m.texture];
			md[i].texRows = c->texRows[m.texture];
			md[i].texCols = c->texCols[m.texture];
		}
		md[i].pad[0] = md[i].pad[1] = md[i].pad[2] = 0;
	}
	if (szMat && !inlineTables) MR_CUDA(c, cudaMemcpyAsync(c->mats.p, md, szMat, cudaMemcpyHostToDevice, c->stream));
	c->h2dBytesLastFrame = szStat + szCB + (inlineTables ? 0 : szDyn + szMat) + sizeof(FrameParams);

	// ---- frame parameters ----
	FrameParams fp;
	memset(&fp, 0, sizeof(fp));
	memcpy(fp.P, f->projection, sizeof(fp.P));
	memcpy(fp.light, f->light, 12);
	fp.ambient = f->ambient;
	memcpy(fp.bg, f->background, 12);
	fp.znear = f->znear;
	fp.wf = (float)c->w;
	fp.hf = (float)c->h;
	fp.w = c->w;
	fp.h = c->h;
	fp.tilesX = c->tilesX;
	fp.tilesY = c->tilesY;
	int rb = 0, re = c->h;
	if (f->row_end > f->row_begin)
	{
		rb = std::max(0, f->row_begin);
		re = std::min(c->h, f->row_end);
	}
	fp.rowBegin = rb;
	fp.rowEnd = re;
	fp.tileRow0 = rb >> MR_TILE_SHIFT;
	fp.tileRows = (re > rb) ? ((re + MR_TILE - 1) >> MR_TILE_SHIFT) - fp.tileRow0 : 0;
	{
		const bool vecOk = (c->w & 7) == 0 && !f->keep;
		fp.fullTx = vecOk ? c->w >> MR_TILE_SHIFT : 0;
		fp.fullTy0 = (rb + MR_TILE - 1) >> MR_TILE_SHIFT;
		fp.fullTy1 = std::min(re, c->h) >> MR_TILE_SHIFT;
		for (int k = 0; k < 12; k++)
			fp.bgPattern[k] = f->background[k % 3];
	}
	fp.persp = f->projection[15] == 0.0f;
	fp.tightScan = (c->noTightScan || (c->debugFlags & 16)) ? 0 : 1;
	{
		const float* P = f->projection;
		fp.stdProj = (fp.persp && P[1] == 0.0f && P[3] == 0.0f && P[4] == 0.0f && P[7] == 0.0f && P[12] == 0.0f && P[13] == 0.0f && P[14] == -1.0f &&
		              f->znear < 0.0f && !c->noStdProj && !(c->debugFlags & 8)) ? 1 : 0;
	}
	{
		const float* P = f->projection;
		const bool form = fp.persp && P[1] == 0.0f && P[3] == 0.0f && P[4] == 0.0f && P[7] == 0.0f && P[12] == 0.0f && P[13] == 0.0f && P[14] == -1.0f &&
		                  P[0] != 0.0f && P[5] != 0.0f;
		fp.unproject = (form && MR_UNPROJECT_POSITIONS && !(c->debugFlags & 2048)) ? 1 : 0;
		if (fp.unproject)
		{
			// ndc.x = (j + 0.5) * 2 / w - 1 = (P0 x + P2 z) / -z  =>  x / -z = (ndc.x + P2) / P0; likewise y with ndc.y = 1 - (i + 0.5) * 2 / h
			fp.unprojX[0] = (float)(2.0 / (double)c->w / (double)P[0]);
			fp.unprojX[1] = (float)(((double)P[2] - 1.0) / (double)P[0]);
			fp.unprojY[0] = (float)(-2.0 / (double)c->h / (double)P[5]);
			fp.unprojY[1] = (float)(((double)P[6] + 1.0) / (double)P[5]);
		}
	}
	{
		// Cluster culling needs the standard perspective form (w_clip = -z_view, x and y not mirrored
		const float* P = f->projection;
		// ... and no translation in clip x / y: the bounds below are planes through the eye)
		const bool standard = fp.persp && P[12] == 0.0f && P[13] == 0.0f && P[14] == -1.0f && P[3] == 0.0f && P[7] == 0.0f &&
		                      (double)P[0] * P[5] - (double)P[1] * P[4] > 0.0;
		fp.cullClusters = (standard && !c->noClusterCull && !(c->debugFlags & 4)) ? 1 : 0;
		if (fp.cullClusters)
		{
			// Columns / rows this frame can touch, widened by 1.5 pixels, as NDC bounds; each bound is a plane
			// through the eye: (x_ndc >= xl) <=> (row0 - xl * row3) . p >= 0 for points in front of the camera.
			const float xl = -1.0f - 3.0f / (float)c->w, xr = 1.0f + 3.0f / (float)c->w;
			const float yt = 1.0f - 2.0f * ((float)rb - 1.5f) / (float)c->h, yb = 1.0f - 2.0f * ((float)re + 1.5f) / (float)c->h;
			const float sgn[4] = { 1.0f, -1.0f, 1.0f, -1.0f }, bound[4] = { xl, xr, yb, yt };
			for (int k = 0; k < 4; k++)
			{
				const float* row = P + 4 * (k / 2);
				double n[4];
				for (int j = 0; j < 4; j++)
					n[j] = sgn[k] * ((double)row[j] - (double)bound[k] * (double)P[12 + j]);
				const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
				if (!(len > 0))
				{
					fp.cullClusters = 0;
					break;
				}
				for (int j = 0; j < 4; j++)
					fp.cullPlanes[k][j] = (float)(n[j] / len);
			}
		}
	}
	fp.lightIsPoint = f->light_is_point;
	fp.lighting = f->lighting;
	fp.texturing = f->texturing;
	fp.saveNormals = f->save_normals;
	fp.keep = f->keep;
	fp.nRenderables = nR;
	fp.nTriInst = c->nTriInst;
	fp.nTriReal = (int)triReal;
	fp.debug = c->debugFlags;
	fp.inlineTables = inlineTables ? 1 : 0;
	if (inlineTables)
	{
		memcpy(fp.rstatInline, rs.data(), sizeof(RStat) * (size_t)nR);
		memcpy(fp.rdynInline, rd, sizeof(RDyn) * (size_t)nR);
		memcpy(fp.matsInline, md, sizeof(MatDev) * (size_t)f->n_materials);
	}
	fp.binCap = c->binCap;
	fp.ovfCap = (int)std::min<size_t>(c->ovfCap, 0x7fffffff);
	fp.meshlets = c->meshlets.as<unsigned char>();
	fp.meshletDir = c->meshletDir.as<MeshletDir>();
	fp.texels = c->texels.as<float4>();
	fp.rstat = c->rstat.as<RStat>();
	fp.rdyn = c->rdyn.as<RDyn>();
	fp.mats = c->mats.as<MatDev>();
	fp.triBlockCl = c->triBlockCl.as<int>();
	fp.clusters = c->clusters.as<float4>();
	fp.geomVertCap = c->geomVertCap;
	fp.visEntries = c->visEntries.as<GeomEntry>();
	fp.geomSync = c->geomSync.as<int>();
	fp.gkeys = c->gkeys.as<unsigned long long>();
	fp.recs = c->recs.as<float4>();
	fp.recs1 = c->recs1.as<float4>();
	fp.tileCount = c->tileCount.as<int2>();
	fp.ovfPairs = c->ovfPairs.as<int2>();
	fp.bins = c->bins.as<int>();
	fp.ctr = c->ctr.as<Counters>() + slotIndex; // per-slot counters: the read-back overlaps the next frame
	fp.ctrNext = c->ctr.as<Counters>() + (slotIndex + 1) % mr_ctx::kSlots;
	fp.image = (c->remoteImage && !f->keep) ? (float*)c->remoteImage : c->imageSlot[c->outCur].as<float>();
	fp.depth = (c->remoteDepth && !f->keep) ? (float*)c->remoteDepth : c->depthSlot[c->outCur].as<float>();
	fp.normals = f->save_normals ? c->normals.as<float>() : 0;
	fp.winner = (c->debugFlags & 1) ? c->winner.as<int>() : 0;

	fp.chkEnable = chkEnable ? 1 : 0;
	fp.chkMinTiles = (c->debugFlags & 128) ? 2 : MR_CHK_MIN_TILES; // (forced on for verification: also for narrow triangles)
	fp.chkPool = c->chkPool.as<float2>();
	fp.chkItems = c->chkItems.as<int4>();
	fp.chkCap = (int)std::min<size_t>(c->chkPool.cap / sizeof(float2), 0x7ff00000);
	fp.chkItemCap = (int)std::min<size_t>(c->chkItems.cap / sizeof(int4), 0x7fffffff);
	fp.sparseStores = (c->sparseRemote && c->remoteImage && !f->keep && !f->save_normals && !(c->debugFlags & 1)) ? 1 : 0;
	mrk_launch_frame(fp, c->geomGrid, c->geomSmem, c->stream, ev, ev ? 0 : c->timingStart, ev ? 0 : c->timingStop, !c->noPdl, c->gateWord, c->gateValue);
	c->gateWord = 0;
	if (!ev)
		c->timingStart = c->timingStop = 0;
	MR_CUDA(c, cudaGetLastError());
	slot.wholeFrame = !f->keep && rb == 0 && re == c->h && !c->remoteImage;
	memcpy(slot.bg, f->background, sizeof(slot.bg));
	MR_CUDA(c, cudaEventRecord(slot.kernelsDone, c->stream));
	MR_CUDA(c, cudaStreamWaitEvent(c->aux, slot.kernelsDone, 0));
	MR_CUDA(c, cudaMemcpyAsync(slot.hostCtr, fp.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, c->aux));
	MR_CUDA(c, cudaEventRecord(slot.done, c->aux));
	slot.pending = true;
	c->slotNewest = slotIndex;
	c->slotNext = (slotIndex + 1) % mr_ctx::kSlots;
	c->stats.kernels_launched = (fp.tileRows > 0) ? (chkEnable ? 3 : 2) : 1;
	return MR_OK;
}

int rememberFrame(mr_ctx* c, const mr_frame* f)
{
	c->lastFrame = *f;
	// (what a re-render after an overflow, mr_profile_frame and the immediate-mode calls go back to; a megabyte per
	// frame for 10 000 renderables, copied in pieces by the host threads when there are that many)
	c->lastRenderables.resize((size_t)f->n_renderables);
	c->lastMaterials.resize((size_t)f->n_materials);
	struct Copy
	{
		mr_ctx* c; const mr_frame* f;
		void operator()(int chunk) const
		{
			const int i0 = chunk * kHostChunk;
			if (i0 < f->n_renderables)
				memcpy(&c->lastRenderables[(size_t)i0], f->renderables + i0, sizeof(mr_renderable) * (size_t)std::min((int)kHostChunk, f->n_renderables - i0));
			if (i0 < f->n_materials)
				memcpy(&c->lastMaterials[(size_t)i0], f->materials + i0, sizeof(mr_material) * (size_t)std::min((int)kHostChunk, f->n_materials - i0));
		}
	} copy = { c, f };
	const int most = std::max(f->n_renderables, f->n_materials);
	minirender::hostpool::parallelFor(most >= kHostPoolMin ? minirender::hostpool::Pool::get() : 0, (most + kHostChunk - 1) / kHostChunk, copy);
	c->lastFrame.renderables = 0;
	c->lastFrame.materials = 0;
	c->haveFrame = true;
	return MR_OK;
}

int readBack(mr_ctx* c, void* host, const void* dev, size_t bytes)
{
	MR_CUDA(c, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
	MR_CUDA(c, cudaStreamSynchronize(c->stream));
	return MR_OK;
}

// Meshlets of one mesh: for every MR_CLUSTER consecutive triangles, the distinct (position, normal, texcoord)
// index triples of their corners in order of first use ("local corners": what paintMesh's loop C gathers,
// Renderer.cpp:351-380, deduplicated), stored as two float4 planes, and the triangles as three 10-bit local
// indices each. Appends the blobs to `blob` (16-byte aligned), fills dir[] and raises maxNv.
void buildMeshlets(const mr_mesh_desc& m, bool hasUV, std::vector<unsigned char>& blob, MeshletDir* dir, int& maxNv)
{
	const int nCl = (m.n_triangles + MR_CLUSTER - 1) / MR_CLUSTER;
	// open-addressing table of the cluster's corners; entries of older clusters are recognised by their stamp
	enum { kTable = 1024 };
	struct Slot { int ip, in, iu, local, stamp; };
	std::vector<Slot> table((size_t)kTable);
	for (int i = 0; i < kTable; i++)
		table[i].stamp = -1;
	float verts[MR_MESHLET_MAX_VERTS][8];
	uint32_t idx[MR_CLUSTER];
	for (int k = 0; k < nCl; k++)
	{
		const int t0 = k * MR_CLUSTER, t1 = std::min(m.n_triangles, t0 + MR_CLUSTER);
		int nv = 0;
		memset(idx, 0, sizeof(idx));
		for (int t = t0; t < t1; t++)
		{
			uint32_t packed = 0;
			for (int j = 0; j < 3; j++)
			{
				const size_t corner = 3 * (size_t)t + j;
				const int ip = m.idx_pos[corner], in = m.idx_nrm[corner], iu = hasUV ? m.idx_uv[corner] : -1;
				uint32_t h = ((uint32_t)ip * 2654435761u) ^ ((uint32_t)in * 40503u) ^ ((uint32_t)iu * 2246822519u);
				h = (h ^ (h >> 15)) & (kTable - 1);
				int local = -1;
				for (;; h = (h + 1) & (kTable - 1))
				{
					Slot& sl = table[h];
					if (sl.stamp != k)
					{
						sl.stamp = k; sl.ip = ip; sl.in = in; sl.iu = iu; sl.local = nv;
						local = nv;
						const float* p = m.positions + 3 * (size_t)ip;
						const float* n = m.normals + 3 * (size_t)in;
						float* o = verts[nv++];
						o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
						o[3] = n[0]; o[4] = n[1]; o[5] = n[2];
						o[6] = hasUV ? m.texcoords[2 * (size_t)iu] : 0.0f;
						o[7] = hasUV ? m.texcoords[2 * (size_t)iu + 1] : 0.0f;
						break;
					}
					if (sl.ip == ip && sl.in == in && sl.iu == iu)
					{
						local = sl.local;
						break;
					}
				}
				packed |= (uint32_t)local << (10 * j);
			}
			idx[t - t0] = packed;
		}
		const size_t at = blob.size(); // a multiple of 16: every blob is
		blob.resize(at + (size_t)MR_MESHLET_BYTES(nv));
		float* p0 = reinterpret_cast<float*>(blob.data() + at);
		float* p1 = p0 + 4 * (size_t)nv;
		for (int v = 0; v < nv; v++)
		{
			p0[4 * v] = verts[v][0]; p0[4 * v + 1] = verts[v][1]; p0[4 * v + 2] = verts[v][2]; p0[4 * v + 3] = verts[v][3];
			p1[4 * v] = verts[v][4]; p1[4 * v + 1] = verts[v][5]; p1[4 * v + 2] = verts[v][6]; p1[4 * v + 3] = verts[v][7];
		}
		memcpy(p1 + 4 * (size_t)nv, idx, sizeof(idx));
		bool finite = true;
		for (int v = 0; v < nv && finite; v++)
			for (int a = 0; a < 3; a++)
				if (!(fabsf(verts[v][a]) <= 3.0e38f)) // NaN or infinite
					finite = false;
		dir[k].off16 = (uint32_t)(at / 16);
		dir[k].nv = (uint32_t)nv | (finite ? 0u : (MR_MESHLET_NONFINITE << 16));
		maxNv = std::max(maxNv, nv);
	}
}

// Cull clusters of one mesh: for every MR_CLUSTER consecutive triangles, a bounding sphere of their
// vertices and a cone around their geometric normals, (b - a) x (c - a). k_setup drops a whole CTA
// of triangles when the sphere lies outside the rows / columns of the frame or nearer than the near
// plane, or when every normal of the cone points away from the eye — only triangles the reference
// itself discards (off-screen reject Renderer.cpp:202, near test :169-177, area cull :205-210),
// decided with generous margins, so the image does not change.
void buildClusters(const mr_mesh_desc& m, float* out /* 8 floats per cluster */)
{
	const double kMargin = 0.07; // radians (4 degrees) added to the cone half-angle
	const int nCl = (m.n_triangles + MR_CLUSTER - 1) / MR_CLUSTER;
	for (int k = 0; k < nCl; k++)
	{
		const int t0 = k * MR_CLUSTER, t1 = std::min(m.n_triangles, t0 + MR_CLUSTER);
		double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 }, sum[3] = { 0, 0, 0 };
		bool finite = true;
		for (int t = t0; t < t1; t++)
		{
			const float* v[3];
			for (int j = 0; j < 3; j++)
			{
				v[j] = m.positions + 3 * (size_t)m.idx_pos[3 * (size_t)t + j];
				for (int a = 0; a < 3; a++)
				{
					const double x = v[j][a];
					if (!(x == x) || x > 1e30 || x < -1e30)
						finite = false;
					lo[a] = std::min(lo[a], x);
					hi[a] = std::max(hi[a], x);
				}
			}
			const double e1[3] = { (double)v[1][0] - v[0][0], (double)v[1][1] - v[0][1], (double)v[1][2] - v[0][2] };
			const double e2[3] = { (double)v[2][0] - v[0][0], (double)v[2][1] - v[0][1], (double)v[2][2] - v[0][2] };
			const double n[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
			const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
			if (len > 0 && len == len) // a zero-area triangle has no facing (and never draws)
				for (int a = 0; a < 3; a++)
					sum[a] += n[a] / len;
		}
		float* o = out + 8 * (size_t)k;
		if (!finite)
		{
			o[0] = o[1] = o[2] = 0.0f; o[3] = INFINITY; // never culled
			o[4] = o[5] = 0.0f; o[6] = 1.0f; o[7] = 2.0f;
			continue;
		}
		const double c[3] = { 0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2]) };
		double r2 = 0;
		for (int t = t0; t < t1; t++)
			for (int j = 0; j < 3; j++)
			{
				const float* v = m.positions + 3 * (size_t)m.idx_pos[3 * (size_t)t + j];
				const double d[3] = { v[0] - c[0], v[1] - c[1], v[2] - c[2] };
				r2 = std::max(r2, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
			}
		o[0] = (float)c[0]; o[1] = (float)c[1]; o[2] = (float)c[2];
		o[3] = (float)(sqrt(r2) * 1.001 + 1e-30) + fabsf((float)c[0]) * 1e-6f + fabsf((float)c[1]) * 1e-6f + fabsf((float)c[2]) * 1e-6f;
		const double sl = sqrt(sum[0] * sum[0] + sum[1] * sum[1] + sum[2] * sum[2]);
		o[4] = o[5] = 0.0f; o[6] = 1.0f; o[7] = 2.0f; // no usable cone unless shown otherwise
		if (sl > 1e-6)
		{
			const double ax[3] = { sum[0] / sl, sum[1] / sl, sum[2] / sl };
			double cosMin = 1.0;
			for (int t = t0; t < t1; t++)
			{
				const float* v0 = m.positions + 3 * (size_t)m.idx_pos[3 * (size_t)t];
				const float* v1 = m.positions + 3 * (size_t)m.idx_pos[3 * (size_t)t + 1];
				const float* v2 = m.positions + 3 * (size_t)m.idx_pos[3 * (size_t)t + 2];
				const double e1[3] = { (double)v1[0] - v0[0], (double)v1[1] - v0[1], (double)v1[2] - v0[2] };
				const double e2[3] = { (double)v2[0] - v0[0], (double)v2[1] - v0[1], (double)v2[2] - v0[2] };
				const double n[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
				const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
				if (len > 0 && len == len)
					cosMin = std::min(cosMin, (n[0] * ax[0] + n[1] * ax[1] + n[2] * ax[2]) / len);
			}
			const double alpha = acos(std::max(-1.0, std::min(1.0, cosMin))) + kMargin;
			if (alpha < 1.5) // below ~86 degrees: a cone that can ever face away as a whole
			{
				o[4] = (float)ax[0]; o[5] = (float)ax[1]; o[6] = (float)ax[2];
				o[7] = (float)sin(alpha);
			}
		}
	}
}

}

extern "C" {

int mr_abi_version(void) { return MR_ABI_VERSION; }

int mr_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess)
	{
		cudaGetLastError();
		return 0;
	}
	return n;
}

mr_ctx* mr_create(int device, int* status)
{
	int st = MR_OK;
	mr_ctx* c = 0;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
	{
		cudaGetLastError();
		st = MR_E_NO_DEVICE;
	}
	else if (device < 0 || device >= n)
		st = MR_E_INVALID;
	else
	{
		c = new mr_ctx();
		c->device = device;
		Bind bind(device);
		c->smCount = 148;
		cudaDeviceGetAttribute(&c->smCount, cudaDevAttrMultiProcessorCount, device);
		c->noClusterCull = getenv("MR_NO_CLUSTER_CULL") != 0; // verification switches, read once
		c->noPdl = getenv("MR_NO_PDL") != 0;
		c->noChain = getenv("MR_NO_CHAIN_CHECKPOINTS") != 0;
		c->noStdProj = getenv("MR_NO_STD_PROJ") != 0;
		c->noTightScan = getenv("MR_NO_TIGHT_SCAN") != 0;
		bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
		c->ownStream = ok;
		ok = ok && cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking) == cudaSuccess;
		ok = ok && cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) == cudaSuccess;
		for (int i = 0; i < 2 && ok; i++)
		{
			ok = ok && cudaEventCreateWithFlags(&c->frameDone[i], cudaEventDisableTiming) == cudaSuccess;
			ok = ok && cudaEventCreateWithFlags(&c->copyDone[i], cudaEventDisableTiming) == cudaSuccess;
		}
		for (int i = 0; i < mr_ctx::kSlots && ok; i++)
		{
			ok = ok && cudaEventCreateWithFlags(&c->slots[i].done, cudaEventDisableTiming) == cudaSuccess;
			ok = ok && cudaEventCreateWithFlags(&c->slots[i].kernelsDone, cudaEventDisableTiming) == cudaSuccess;
			ok = ok && cudaMallocHost((void**)&c->slots[i].hostCtr, sizeof(Counters)) == cudaSuccess;
			if (ok)
				memset(c->slots[i].hostCtr, 0, sizeof(Counters));
		}
		if (ok)
		{
			const int fma = mrk_selftest_no_fma(c->stream);
			if (fma != 0)
			{
				fprintf(stderr, "minirender_b200: FMA contraction self-test failed (%d); build with --fmad=false\n", fma);
				ok = false;
			}
		}
		if (!ok)
		{
			cudaGetLastError();
			mr_destroy(c);
			c = 0;
			st = MR_E_CUDA;
		}
	}
	if (status)
		*status = st;
	return c;
}

void mr_destroy(mr_ctx* c)
{
	if (!c)
		return;
	Bind bind(c->device);
	if (c->stream)
		cudaStreamSynchronize(c->stream);
	DevBuf* bufs[] = { &c->meshlets, &c->meshletDir, &c->texels, &c->clusters, &c->triBlockCl, &c->visEntries, &c->geomSync, &c->rstat,
		               &c->rdyn, &c->mats, &c->recs, &c->tileCount,
		               &c->ovfPairs, &c->chkPool, &c->chkItems, &c->bins, &c->recs1, &c->gkeys, &c->ctr, &c->imageSlot[0], &c->depthSlot[0], &c->imageSlot[1], &c->depthSlot[1], &c->normals, &c->winner, &c->scratchOut, &c->flushBuf, &c->syncWords };
	for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); i++)
		bufs[i]->release();
	for (int i = 0; i < mr_ctx::kSlots; i++)
	{
		c->slots[i].stage.release();
		if (c->slots[i].hostCtr)
			cudaFreeHost(c->slots[i].hostCtr);
		if (c->slots[i].done)
			cudaEventDestroy(c->slots[i].done);
		if (c->slots[i].kernelsDone)
			cudaEventDestroy(c->slots[i].kernelsDone);
	}
	if (c->ownStream && c->stream)
		cudaStreamDestroy(c->stream);
	if (c->aux)
	{
		cudaStreamSynchronize(c->aux);
		cudaStreamDestroy(c->aux);
	}
	if (c->copy)
	{
		cudaStreamSynchronize(c->copy);
		cudaStreamDestroy(c->copy);
	}
	for (int i = 0; i < 2; i++)
	{
		if (c->frameDone[i]) cudaEventDestroy(c->frameDone[i]);
		if (c->copyDone[i]) cudaEventDestroy(c->copyDone[i]);
	}
	delete c;
}

const char* mr_last_error(const mr_ctx* c) { return c ? c->error.c_str() : "no context (no CUDA device?)"; }

int mr_set_stream(mr_ctx* c, void* s)
{
	if (!c)
		return MR_E_INVALID;
	Bind bind(c->device);
	MR_CUDA(c, cudaStreamSynchronize(c->stream));
	if (c->ownStream)
	{
		cudaStreamDestroy(c->stream);
		c->ownStream = false;
	}
	if (s)
		c->stream = (cudaStream_t)s;
	else
	{
		MR_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		c->ownStream = true;
	}
	return MR_OK;
}

int mr_set_timing_events(mr_ctx* c, void* start, void* stop)
{
	if (!c)
		return MR_E_INVALID;
	c->timingStart = (cudaEvent_t)start;
	c->timingStop = (cudaEvent_t)stop;
	return MR_OK;
}

int mr_set_debug(mr_ctx* c, int flags)
{
	if (!c)
		return MR_E_INVALID;
	c->debugFlags = flags;
	return MR_OK;
}

int mr_set_size(mr_ctx* c, int w, int h)
{
	if (!c)
		return MR_E_INVALID;
	if (w <= 0 || h <= 0 || w > 32768 || h > 32768)
		return setError(c, MR_E_INVALID, "image size %dx%d out of range (1..32768)", w, h);
	Bind bind(c->device);
	if (w == c->w && h == c->h)
		return MR_OK;
	{
		int rc = finishFrame(c);
		if (rc)
			return rc;
	}
	c->w = w;
	c->h = h;
	c->tilesX = (w + MR_TILE - 1) / MR_TILE;
	c->tilesY = (h + MR_TILE - 1) / MR_TILE;
	c->binCap = 0;
	c->bins.release();
	for (int sl = 0; sl < 2; sl++)
	{
		c->imageSlot[sl].release();
		c->depthSlot[sl].release();
	}
	c->gkeys.release();
	c->normals.release();
	c->winner.release();
	c->haveFrame = false;
	return ensureOutputs(c, false, false);
}

int mr_upload_scene(mr_ctx* c, const mr_scene_desc* s)
{
	if (!c || !s || s->n_meshes < 0 || s->n_textures < 0)
		return setError(c, MR_E_INVALID, "bad scene descriptor");
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	// ---- validate and lay out ----
	std::vector<MeshDev> md((size_t)s->n_meshes);
	long long nTri = 0;
	size_t nClusters = 0;
	for (int i = 0; i < s->n_meshes; i++)
	{
		const mr_mesh_desc& m = s->meshes[i];
		if (m.n_positions < 0 || m.n_normals < 0 || m.n_texcoords < 0 || m.n_triangles < 0)
			return setError(c, MR_E_INVALID, "mesh %d: negative count", i);
		if (m.n_triangles > 0 && (!m.idx_pos || !m.idx_nrm || !m.positions || !m.normals))
			return setError(c, MR_E_INVALID, "mesh %d: missing positions/normals/index arrays", i);
		const bool hasUV = m.n_texcoords > 0 && m.idx_uv != 0 && m.texcoords != 0;
		for (long long k = 0; k < 3LL * m.n_triangles; k++)
		{
			if ((unsigned)m.idx_pos[k] >= (unsigned)m.n_positions)
				return setError(c, MR_E_INVALID, "mesh %d: position index %d out of range at corner %lld", i, m.idx_pos[k], k);
			if ((unsigned)m.idx_nrm[k] >= (unsigned)m.n_normals)
				return setError(c, MR_E_INVALID, "mesh %d: normal index %d out of range at corner %lld", i, m.idx_nrm[k], k);
			if (hasUV && (unsigned)m.idx_uv[k] >= (unsigned)m.n_texcoords)
				return setError(c, MR_E_INVALID, "mesh %d: texcoord index %d out of range at corner %lld", i, m.idx_uv[k], k);
		}
		md[i].clusterBase = (int)nClusters;
		md[i].nTri = m.n_triangles;
		md[i].hasUV = hasUV;
		md[i].pad = 0;
		nTri += m.n_triangles;
		nClusters += (size_t)(m.n_triangles + MR_CLUSTER - 1) / MR_CLUSTER;
	}
	if (nTri > 0x2aaaaaaaLL)
		return setError(c, MR_E_INVALID, "scene too large for 32-bit indexing");
	long long nTexel = 0;
	std::vector<int> to((size_t)s->n_textures), tr((size_t)s->n_textures), tc((size_t)s->n_textures);
	for (int i = 0; i < s->n_textures; i++)
	{
		const mr_texture_desc& t = s->textures[i];
		if (t.rows <= 0 || t.cols <= 0 || !t.texels)
			return setError(c, MR_E_INVALID, "texture %d: empty", i);
		to[i] = (int)nTexel;
		tr[i] = t.rows;
		tc[i] = t.cols;
		nTexel += (long long)t.rows * t.cols;
	}
	if (nTexel > 0x3fffffffLL)
		return setError(c, MR_E_INVALID, "textures too large");

	// ---- meshlets + cull clusters (host pass over the triangles, once per upload) ----
	std::vector<unsigned char> blob;
	std::vector<MeshletDir> dir(std::max<size_t>(nClusters, 1));
	std::vector<float> clusters(8 * std::max<size_t>(nClusters, 1), 0.0f);
	int maxNv = 4;
	for (int i = 0; i < s->n_meshes; i++)
		if (s->meshes[i].n_triangles > 0)
		{
			buildMeshlets(s->meshes[i], md[i].hasUV != 0, blob, dir.data() + md[i].clusterBase, maxNv);
			buildClusters(s->meshes[i], clusters.data() + 8 * (size_t)md[i].clusterBase);
		}
	if (blob.size() / 16 > 0xffffffffULL)
		return setError(c, MR_E_INVALID, "scene too large: %zu bytes of meshlets", blob.size());
	const int nvCap = (maxNv + 3) & ~3;
	int grid = 0, smem = 0;
	if (mrk_geom_config(nvCap, c->smCount, &grid, &smem) != 0)
	{
		cudaGetLastError();
		return setError(c, MR_E_CUDA, "k_geom does not fit this device (meshlets of %d corners)", nvCap);
	}
	c->geomVertCap = nvCap;
	c->geomGrid = grid;
	c->geomSmem = smem;

	MR_CUDA(c, c->meshlets.ensure(std::max<size_t>(blob.size(), 16)));
	MR_CUDA(c, c->meshletDir.ensure(sizeof(MeshletDir) * dir.size()));
	MR_CUDA(c, c->clusters.ensure(sizeof(float) * clusters.size()));
	MR_CUDA(c, c->texels.ensure(sizeof(float4) * (size_t)std::max(nTexel, 1LL)));
	if (!blob.empty())
		MR_CUDA(c, cudaMemcpyAsync(c->meshlets.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, c->stream));
	MR_CUDA(c, cudaMemcpyAsync(c->meshletDir.p, dir.data(), sizeof(MeshletDir) * dir.size(), cudaMemcpyHostToDevice, c->stream));
	MR_CUDA(c, cudaMemcpyAsync(c->clusters.p, clusters.data(), sizeof(float) * clusters.size(), cudaMemcpyHostToDevice, c->stream));

	// packed rgb texel arrays go through a device staging buffer and are widened to float4 there
	size_t maxRaw = 0;
	for (int i = 0; i < s->n_textures; i++)
		maxRaw = std::max(maxRaw, sizeof(float) * 3 * (size_t)s->textures[i].rows * s->textures[i].cols);
	MR_CUDA(c, c->scratchOut.ensure(std::max<size_t>(maxRaw, 16)));
	float* raw = c->scratchOut.as<float>();
	for (int i = 0; i < s->n_textures; i++)
	{
		const mr_texture_desc& t = s->textures[i];
		const size_t n = (size_t)t.rows * t.cols;
		MR_CUDA(c, cudaMemcpyAsync(raw, t.texels, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
		mrk_launch_pack(c->texels.as<float4>() + to[i], raw, (int)n, 3, 0.0f, c->stream);
	}
	MR_CUDA(c, cudaGetLastError());
	MR_CUDA(c, cudaStreamSynchronize(c->stream)); // host arrays are only borrowed for this call
	c->hostMeshes.swap(md);
	c->texOffset.swap(to);
	c->texRows.swap(tr);
	c->texCols.swap(tc);
	c->haveScene = true;
	c->sceneSerial++;
	c->haveFrame = false;
	return MR_OK;
}

int mr_render(mr_ctx* c, const mr_frame* f)
{
	if (!c || !f)
		return MR_E_INVALID;
	if (!c->haveScene)
		return setError(c, MR_E_NO_SCENE, "mr_render before mr_upload_scene");
	if (c->w <= 0)
		return setError(c, MR_E_INVALID, "mr_render before mr_set_size");
	if (f->n_renderables < 0 || f->n_materials < 0 || (f->n_renderables > 0 && !f->renderables) || (f->n_materials > 0 && !f->materials))
		return setError(c, MR_E_INVALID, "bad frame descriptor");
	Bind bind(c->device);
	rememberFrame(c, f);
	return launchFrame(c, f, 0);
}

// Delivers frame `index` of a batch to the sink once it is complete. `slot` / `out` are the frame slot (counters) and
// the output set it was rendered into. A frame that overflowed its spill list is rendered again first (everything in
// flight is finished for that: rare, and the capacities have been raised by then).
static int deliverBatchFrame(mr_ctx* c, const mr_frame* frames, int index, int slot, int out, mr_frame_sink sink, void* user)
{
	int rc = retireSlot(c, slot); // waits for the frame's kernels and its counters, not for younger frames
	if (rc < 0)
		return rc;
	if (rc > 0)
	{
		rc = finishFrame(c); // (re-renders the newest frame too if that one overflowed as well)
		if (rc)
			return rc;
		const int keepCur = c->outCur;
		c->outCur = out; // render into the set the frame was in: the younger frame's set stays untouched
		MR_CUDA(c, cudaMemsetAsync(c->gkeys.p, 0xff, (size_t)c->w * c->h * 8, c->stream));
		rc = launchFrame(c, &frames[index], 0, true);
		if (rc == MR_OK)
			rc = retireSlot(c, c->slotNewest) > 0 ? setError(c, MR_E_OVERFLOW, "batch frame %d kept overflowing its spill list", index) : MR_OK;
		c->outCur = keepCur;
		if (rc)
			return rc;
		MR_CUDA(c, cudaMemsetAsync(c->gkeys.p, 0xff, (size_t)c->w * c->h * 8, c->stream));
		MR_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	sink(user, index, c->imageSlot[out].as<float>(), c->depthSlot[out].as<float>());
	return MR_OK;
}

int mr_render_batch(mr_ctx* c, int n, const mr_frame* frames, mr_frame_sink sink, void* user)
{
	if (!c || n < 0 || (n > 0 && !frames))
		return MR_E_INVALID;
	Bind bind(c->device);
	// With two output sets the kernels of frame i + 1 are in flight while the sink looks at frame i; with one set
	// every frame is finished and delivered before the next one is launched.
	int prevSlot = -1, prevOut = -1;
	for (int i = 0; i < n; i++)
	{
		if (sink && prevSlot >= 0 && c->outSlots < 2)
		{
			const int rc = deliverBatchFrame(c, frames, i - 1, prevSlot, prevOut, sink, user);
			if (rc)
				return rc;
			prevSlot = -1;
		}
		int rc = mr_render(c, &frames[i]);
		if (rc)
			return rc;
		const int slot = c->slotNewest, out = c->outCur;
		if (sink && prevSlot >= 0)
		{
			rc = deliverBatchFrame(c, frames, i - 1, prevSlot, prevOut, sink, user);
			if (rc)
				return rc;
		}
		prevSlot = slot;
		prevOut = out;
	}
	if (sink && prevSlot >= 0)
	{
		const int rc = deliverBatchFrame(c, frames, n - 1, prevSlot, prevOut, sink, user);
		if (rc)
			return rc;
	}
	return sink ? MR_OK : finishFrame(c);
}

int mr_download(mr_ctx* c, void* host, const void* d_src, size_t bytes)
{
	if (!c || !host || !d_src)
		return MR_E_INVALID;
	Bind bind(c->device);
	MR_CUDA(c, cudaMemcpy(host, d_src, bytes, cudaMemcpyDeviceToHost));
	return MR_OK;
}

int mr_synchronize(mr_ctx* c)
{
	if (!c)
		return MR_E_INVALID;
	Bind bind(c->device);
	return finishFrame(c);
}

int mr_read_image(mr_ctx* c, float* host)
{
	if (!c || !host)
		return MR_E_INVALID;
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	return readBack(c, host, c->imageSlot[c->outCur].p, (size_t)c->w * c->h * 12);
}

int mr_read_depth(mr_ctx* c, float* host)
{
	if (!c || !host)
		return MR_E_INVALID;
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	return readBack(c, host, c->depthSlot[c->outCur].p, (size_t)c->w * c->h * 4);
}

int mr_read_normals(mr_ctx* c, float* host)
{
	if (!c || !host)
		return MR_E_INVALID;
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	rc = ensureOutputs(c, true, false);
	if (rc)
		return rc;
	return readBack(c, host, c->normals.p, (size_t)c->w * c->h * 12);
}

int mr_read_winner_ids(mr_ctx* c, int32_t* host)
{
	if (!c || !host)
		return MR_E_INVALID;
	if (!(c->debugFlags & 1) || !c->winner.p)
		return setError(c, MR_E_INVALID, "winner ids not enabled (mr_set_debug(ctx, 1) before mr_render)");
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	return readBack(c, host, c->winner.p, (size_t)c->w * c->h * 4);
}

int mr_read_range(mr_ctx* c, const float* P, float znear, float* host)
{
	(void)znear;
	if (!c || !host || !P)
		return MR_E_INVALID;
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	const size_t bytes = (size_t)c->w * c->h * 12;
	MR_CUDA(c, c->scratchOut.ensure(bytes));
	mrk_launch_range(c->depthSlot[c->outCur].as<float>(), c->scratchOut.as<float>(), c->w, c->h, P, c->stream);
	MR_CUDA(c, cudaGetLastError());
	return readBack(c, host, c->scratchOut.p, bytes);
}

int mr_read_rgb8(mr_ctx* c, uint8_t* host)
{
	if (!c || !host)
		return MR_E_INVALID;
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	const size_t n = (size_t)c->w * c->h * 3;
	MR_CUDA(c, c->scratchOut.ensure(n));
	mrk_launch_rgb8(c->imageSlot[c->outCur].as<float>(), c->scratchOut.as<uint8_t>(), n, c->stream);
	MR_CUDA(c, cudaGetLastError());
	return readBack(c, host, c->scratchOut.p, n);
}

int mr_set_output_slots(mr_ctx* c, int n)
{
	if (!c || n < 1 || n > 2)
		return MR_E_INVALID;
	Bind bind(c->device);
	int rc = finishFrame(c);
	if (rc)
		return rc;
	MR_CUDA(c, cudaStreamSynchronize(c->copy));
	c->copyPending[0] = c->copyPending[1] = false;
	if (n == 1 && c->outCur == 1)
	{
		// keep the newest frame addressable as slot 0
		std::swap(c->imageSlot[0], c->imageSlot[1]);
		std::swap(c->depthSlot[0], c->depthSlot[1]);
	}
	c->outSlots = n;
	c->outCur = 0;
	return c->w > 0 ? ensureOutputs(c, false, false) : MR_OK;
}

int mr_read_image_begin(mr_ctx* c, float* host, int* ticket)
{
	if (!c || !host)
		return MR_E_INVALID;
	Bind bind(c->device);
	const int sl = c->outCur;
	MR_CUDA(c, cudaEventRecord(c->frameDone[sl], c->stream));
	MR_CUDA(c, cudaStreamWaitEvent(c->copy, c->frameDone[sl], 0));
	MR_CUDA(c, cudaMemcpyAsync(host, c->imageSlot[sl].p, (size_t)c->w * c->h * 12, cudaMemcpyDeviceToHost, c->copy));
	c->stats.d2h_bytes = (int64_t)c->w * c->h * 12;
	for (size_t i = 0; i < c->mirrors.size(); i++)
		if (c->mirrors[i].ptr == host)
			c->mirrors[i].w = -1; // (its contents are no longer what the dirty-rectangle bookkeeping assumes)
	MR_CUDA(c, cudaEventRecord(c->copyDone[sl], c->copy));
	c->copyPending[sl] = true;
	c->copyRing[sl] = c->slotNewest;
	if (ticket)
		*ticket = sl;
	return MR_OK;
}

int mr_read_image_dirty_begin(mr_ctx* c, float* host, int* ticket)
{
	if (!c || !host)
		return MR_E_INVALID;
	Bind bind(c->device);
	const int sl = c->outCur;
	if (c->slotNewest < 0)
		return setError(c, MR_E_INVALID, "mr_read_image_dirty_begin before mr_render");
	// the frame's dirty rectangle comes with its counters: wait for them (the frame itself is then complete)
	mr_ctx::Slot& fs = c->slots[c->slotNewest];
	const int rc = retireSlot(c, c->slotNewest);
	if (rc < 0)
		return rc;
	if (rc > 0)
		return setError(c, MR_E_OVERFLOW, "the frame overflowed its spill list (capacities raised): render it again");
	mr_ctx::HostMirror* hm = 0;
	for (size_t i = 0; i < c->mirrors.size(); i++)
		if (c->mirrors[i].ptr == host)
			hm = &c->mirrors[i];
	int x0 = 0, x1 = c->w, y0 = 0, y1 = c->h;
	const bool known = hm && fs.wholeFrame && hm->w == c->w && hm->h == c->h && memcmp(hm->bg, fs.bg, sizeof(fs.bg)) == 0;
	if (known)
	{
		// everything outside (old rectangle U new rectangle) is background in the buffer and in the frame
		const bool oldEmpty = hm->rect[1] <= hm->rect[0] || hm->rect[3] <= hm->rect[2];
		const bool newEmpty = fs.dirty[1] <= fs.dirty[0] || fs.dirty[3] <= fs.dirty[2];
		if (oldEmpty && newEmpty)
			x0 = x1 = y0 = y1 = 0;
		else if (oldEmpty)
		{
			x0 = fs.dirty[0]; x1 = fs.dirty[1]; y0 = fs.dirty[2]; y1 = fs.dirty[3];
		}
		else if (newEmpty)
		{
			x0 = hm->rect[0]; x1 = hm->rect[1]; y0 = hm->rect[2]; y1 = hm->rect[3];
		}
		else
		{
			x0 = std::min(hm->rect[0], fs.dirty[0]); x1 = std::max(hm->rect[1], fs.dirty[1]);
			y0 = std::min(hm->rect[2], fs.dirty[2]); y1 = std::max(hm->rect[3], fs.dirty[3]);
		}
	}
	if (!hm)
	{
		if (c->mirrors.size() >= 16)
			c->mirrors.erase(c->mirrors.begin());
		c->mirrors.push_back(mr_ctx::HostMirror());
		hm = &c->mirrors.back();
		hm->ptr = host;
	}
	hm->w = c->w;
	hm->h = c->h;
	memcpy(hm->bg, fs.bg, sizeof(fs.bg));
	if (fs.wholeFrame)
		memcpy(hm->rect, fs.dirty, sizeof(fs.dirty));
	else
	{
		hm->rect[0] = 0; hm->rect[1] = c->w; hm->rect[2] = 0; hm->rect[3] = c->h; // unknown contents: the next copy is a full one
		hm->w = -1;
	}
	MR_CUDA(c, cudaEventRecord(c->frameDone[sl], c->stream));
	MR_CUDA(c, cudaStreamWaitEvent(c->copy, c->frameDone[sl], 0));
	size_t bytes = 0;
	if (x1 > x0 && y1 > y0)
	{
		const size_t pitch = (size_t)c->w * 12, off = (size_t)y0 * pitch + (size_t)x0 * 12;
		bytes = (size_t)(x1 - x0) * 12 * (size_t)(y1 - y0);
		if (x0 == 0 && x1 == c->w)
			MR_CUDA(c, cudaMemcpyAsync((char*)host + off, (const char*)c->imageSlot[sl].p + off, bytes, cudaMemcpyDeviceToHost, c->copy));
		else
			MR_CUDA(c, cudaMemcpy2DAsync((char*)host + off, pitch, (const char*)c->imageSlot[sl].p + off, pitch, (size_t)(x1 - x0) * 12, (size_t)(y1 - y0),
			                             cudaMemcpyDeviceToHost, c->copy));
	}
	c->stats.d2h_bytes = (int64_t)bytes;
	MR_CUDA(c, cudaEventRecord(c->copyDone[sl], c->copy));
	c->copyPending[sl] = true;
	c->copyRing[sl] = -1; // (already retired)
	if (ticket)
		*ticket = sl;
	return MR_OK;
}

int mr_read_image_dirty_forget(mr_ctx* c, const float* host)
{
	if (!c)
		return MR_E_INVALID;
	for (size_t i = 0; i < c->mirrors.size();)
		if (!host || c->mirrors[i].ptr == host)
			c->mirrors.erase(c->mirrors.begin() + i);
		else
			i++;
	return MR_OK;
}

int mr_read_wait(mr_ctx* c, int ticket)
{
	if (!c || ticket < 0 || ticket > 1)
		return MR_E_INVALID;
	Bind bind(c->device);
	MR_CUDA(c, cudaEventSynchronize(c->copyDone[ticket]));
	// The copied frame's counters follow its kernels on the side stream: a frame that overflowed its spill list
	// is incomplete and must not be taken for good (the capacities have been raised: render it again).
	bool overflowed = c->slotOverflowed;
	c->slotOverflowed = false;
	if (c->copyRing[ticket] >= 0)
	{
		const int rc = retireSlot(c, c->copyRing[ticket]);
		c->copyRing[ticket] = -1;
		if (rc < 0)
			return rc;
		overflowed = overflowed || rc > 0;
	}
	if (overflowed)
		return setError(c, MR_E_OVERFLOW, "a pipelined frame overflowed its spill list (capacities raised): render it again");
	return MR_OK;
}

int mr_read_image_async(mr_ctx* c, float* host)
{
	if (!c || !host)
		return MR_E_INVALID;
	Bind bind(c->device);
	MR_CUDA(c, cudaMemcpyAsync(host, c->imageSlot[c->outCur].p, (size_t)c->w * c->h * 12, cudaMemcpyDeviceToHost, c->stream));
	return MR_OK;
}

int mr_read_rows_async(mr_ctx* c, float* host_rgb, float* host_depth, int rb, int re)
{
	if (!c || rb < 0 || re > c->h || re < rb)
		return MR_E_INVALID;
	Bind bind(c->device);
	const size_t w = (size_t)c->w;
	if (host_rgb)
		MR_CUDA(c, cudaMemcpyAsync(host_rgb + 3 * w * rb, c->imageSlot[c->outCur].as<float>() + 3 * w * rb, 12 * w * (re - rb), cudaMemcpyDeviceToHost, c->stream));
	if (host_depth)
		MR_CUDA(c, cudaMemcpyAsync(host_depth + w * rb, c->depthSlot[c->outCur].as<float>() + w * rb, 4 * w * (re - rb), cudaMemcpyDeviceToHost, c->stream));
	return MR_OK;
}

int mr_host_register(void* host, size_t bytes)
{
	return cudaHostRegister(host, bytes, cudaHostRegisterDefault) == cudaSuccess ? MR_OK : (cudaGetLastError(), MR_E_CUDA);
}

int mr_host_unregister(void* host)
{
	return cudaHostUnregister(host) == cudaSuccess ? MR_OK : (cudaGetLastError(), MR_E_CUDA);
}

int mr_device_buffers(mr_ctx* c, void** image, void** depth, void** normals)
{
	if (!c)
		return MR_E_INVALID;
	if (image) *image = c->imageSlot[c->outCur].p;
	if (depth) *depth = c->depthSlot[c->outCur].p;
	if (normals) *normals = c->normals.p;
	return MR_OK;
}

int mr_write_rows(mr_ctx* c, const void* img, const void* dep, int rb, int re)
{
	if (!c || rb < 0 || re > c->h || re < rb)
		return MR_E_INVALID;
	Bind bind(c->device);
	const size_t w = (size_t)c->w;
	if (img)
		MR_CUDA(c, cudaMemcpyAsync(c->imageSlot[c->outCur].as<float>() + 3 * w * rb, img, 12 * w * (re - rb), cudaMemcpyDeviceToDevice, c->stream));
	if (dep)
		MR_CUDA(c, cudaMemcpyAsync(c->depthSlot[c->outCur].as<float>() + w * rb, dep, 4 * w * (re - rb), cudaMemcpyDeviceToDevice, c->stream));
	return MR_OK;
}

int mr_set_remote_target(mr_ctx* c, void* img, void* dep)
{
	if (!c)
		return MR_E_INVALID;
	c->remoteImage = img;
	c->remoteDepth = dep;
	return MR_OK;
}

/* ---- CUDA IPC: lets another rank's tile rasterizer store straight into this framebuffer ---- */
int mr_ipc_export(mr_ctx* c, void* handles128)
{
	if (!c || !handles128 || !c->imageSlot[c->outCur].p || !c->depthSlot[c->outCur].p)
		return MR_E_INVALID;
	Bind bind(c->device);
	cudaIpcMemHandle_t h[2];
	MR_CUDA(c, cudaIpcGetMemHandle(&h[0], c->imageSlot[c->outCur].p));
	MR_CUDA(c, cudaIpcGetMemHandle(&h[1], c->depthSlot[c->outCur].p));
	memcpy(handles128, h, sizeof(h));
	return MR_OK;
}

int mr_ipc_export_slot(mr_ctx* c, int slot, void* handles128)
{
	if (!c || !handles128 || slot < 0 || slot >= c->outSlots || !c->imageSlot[slot].p || !c->depthSlot[slot].p)
		return MR_E_INVALID;
	Bind bind(c->device);
	cudaIpcMemHandle_t h[2];
	MR_CUDA(c, cudaIpcGetMemHandle(&h[0], c->imageSlot[slot].p));
	MR_CUDA(c, cudaIpcGetMemHandle(&h[1], c->depthSlot[slot].p));
	memcpy(handles128, h, sizeof(h));
	return MR_OK;
}

int mr_output_slot(mr_ctx* c) { return c ? c->outCur : MR_E_INVALID; }

int mr_ipc_open(mr_ctx* c, const void* handles128, void** image, void** depth)
{
	if (!c || !handles128 || !image || !depth)
		return MR_E_INVALID;
	Bind bind(c->device);
	cudaIpcMemHandle_t h[2];
	memcpy(h, handles128, sizeof(h));
	MR_CUDA(c, cudaIpcOpenMemHandle(image, h[0], cudaIpcMemLazyEnablePeerAccess));
	MR_CUDA(c, cudaIpcOpenMemHandle(depth, h[1], cudaIpcMemLazyEnablePeerAccess));
	return MR_OK;
}

int mr_ipc_close(mr_ctx* c, void* image, void* depth)
{
	if (!c)
		return MR_E_INVALID;
	Bind bind(c->device);
	MR_CUDA(c, cudaStreamSynchronize(c->stream));
	if (image) MR_CUDA(c, cudaIpcCloseMemHandle(image));
	if (depth) MR_CUDA(c, cudaIpcCloseMemHandle(depth));
	return MR_OK;
}

int mr_sync_words(mr_ctx* c, int n, void** words)
{
	if (!c || n <= 0 || n > 1024 || !words)
		return MR_E_INVALID;
	Bind bind(c->device);
	if (!c->syncWords.p)
	{
		// (its own allocation: exported to other processes as a whole)
		MR_CUDA(c, c->syncWords.ensure(4096, true));
		MR_CUDA(c, cudaMemsetAsync(c->syncWords.p, 0, 4096, c->stream));
		MR_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	*words = c->syncWords.p;
	return MR_OK;
}

int mr_ipc_export_ptr(mr_ctx* c, const void* dptr, void* handle64)
{
	if (!c || !dptr || !handle64)
		return MR_E_INVALID;
	Bind bind(c->device);
	cudaIpcMemHandle_t h;
	MR_CUDA(c, cudaIpcGetMemHandle(&h, (void*)dptr));
	memcpy(handle64, &h, sizeof(h));
	return MR_OK;
}

int mr_ipc_open_ptr(mr_ctx* c, const void* handle64, void** dptr)
{
	if (!c || !handle64 || !dptr)
		return MR_E_INVALID;
	Bind bind(c->device);
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof(h));
	MR_CUDA(c, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
	return MR_OK;
}

int mr_ipc_close_ptr(mr_ctx* c, void* dptr)
{
	if (!c || !dptr)
		return MR_E_INVALID;
	Bind bind(c->device);
	MR_CUDA(c, cudaStreamSynchronize(c->stream));
	MR_CUDA(c, cudaIpcCloseMemHandle(dptr));
	return MR_OK;
}

int mr_stream_signal(mr_ctx* c, void* word, uint32_t value)
{
	if (!c || !word)
		return MR_E_INVALID;
	Bind bind(c->device);
	mrk_launch_signal((unsigned*)word, value, c->stream);
	MR_CUDA(c, cudaGetLastError());
	return MR_OK;
}

int mr_stream_wait(mr_ctx* c, const void* words, int n, uint32_t value)
{
	if (!c || !words || n <= 0)
		return MR_E_INVALID;
	Bind bind(c->device);
	mrk_launch_wait((const unsigned*)words, n, value, c->stream);
	MR_CUDA(c, cudaGetLastError());
	return MR_OK;
}

int mr_set_raster_gate(mr_ctx* c, const void* word, uint32_t value)
{
	if (!c)
		return MR_E_INVALID;
	c->gateWord = (const unsigned*)word;
	c->gateValue = value;
	return MR_OK;
}

int mr_set_sparse_remote_stores(mr_ctx* c, int flag)
{
	if (!c)
		return MR_E_INVALID;
	c->sparseRemote = flag != 0;
	return MR_OK;
}

int mr_clear_rows(mr_ctx* c, const float* bg, int rb, int re)
{
	if (!c || !bg || rb < 0 || re > c->h || re < rb)
		return MR_E_INVALID;
	Bind bind(c->device);
	if (!c->imageSlot[c->outCur].p)
		return setError(c, MR_E_INVALID, "mr_clear_rows before mr_set_size");
	mrk_launch_clear_rows(c->imageSlot[c->outCur].as<float>(), c->depthSlot[c->outCur].as<float>(), c->w, rb, re, bg[0], bg[1], bg[2], c->stream);
	MR_CUDA(c, cudaGetLastError());
	return MR_OK;
}

int mr_clear_rows_slot(mr_ctx* c, int slot, const float* bg, int rb, int re)
{
	if (!c || !bg || rb < 0 || re > c->h || re < rb || slot < 0 || slot >= c->outSlots)
		return MR_E_INVALID;
	Bind bind(c->device);
	if (!c->imageSlot[slot].p)
		return setError(c, MR_E_INVALID, "mr_clear_rows_slot before mr_set_size");
	mrk_launch_clear_rows(c->imageSlot[slot].as<float>(), c->depthSlot[slot].as<float>(), c->w, rb, re, bg[0], bg[1], bg[2], c->stream);
	MR_CUDA(c, cudaGetLastError());
	return MR_OK;
}

int mr_get_stats(mr_ctx* c, mr_stats* out)
{
	if (!c || !out)
		return MR_E_INVALID;
	c->stats.h2d_bytes = (int64_t)c->h2dBytesLastFrame;
	*out = c->stats;
	return MR_OK;
}

int mr_flush_l2(mr_ctx* c)
{
	if (!c)
		return MR_E_INVALID;
	Bind bind(c->device);
	// Write 256 MiB (twice the 126 MB L2), then stream-read another 256 MiB: the write evicts
	// everything, the read replaces the dirty flush lines with clean ones so that their write-back
	// is not charged to whatever runs next.
	const size_t bytes = (size_t)256 << 20;
	MR_CUDA(c, c->flushBuf.ensure(2 * bytes + 256, true));
	static int v = 0;
	MR_CUDA(c, cudaMemsetAsync(c->flushBuf.p, (++v) & 0xff, bytes, c->stream));
	mrk_launch_flush_read((const char*)c->flushBuf.p + bytes, bytes, (float*)((char*)c->flushBuf.p + 2 * bytes), c->stream);
	MR_CUDA(c, cudaGetLastError());
	return MR_OK;
}

int mr_profile_frame(mr_ctx* c, const mr_frame* f, int repeats)
{
	if (!c || !f || repeats <= 0)
		return MR_E_INVALID;
	Bind bind(c->device);
	int rc = mr_render(c, f); // warm, also settles queue sizes
	if (rc)
		return rc;
	rc = finishFrame(c);
	if (rc)
		return rc;
	cudaEvent_t ev[3];
	for (int i = 0; i < 3; i++)
		MR_CUDA(c, cudaEventCreate(&ev[i]));
	float acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	for (int it = 0; it < repeats; it++)
	{
		if (c->debugFlags & 2)
			mr_flush_l2(c);
		rc = launchFrame(c, f, ev);
		if (rc)
			break;
		MR_CUDA(c, cudaStreamSynchronize(c->stream));
		float ms = 0;
		cudaEventElapsedTime(&ms, ev[0], ev[1]);
		acc[1] += ms; // k_geom
		cudaEventElapsedTime(&ms, ev[1], ev[2]);
		acc[4] += ms; // k_raster
		cudaEventElapsedTime(&ms, ev[0], ev[2]);
		acc[5] += ms;
	}
	for (int i = 0; i < 3; i++)
		cudaEventDestroy(ev[i]);
	for (int i = 0; i < 8; i++)
		c->stats.ms_kernel[i] = acc[i] / repeats;
	if (rc)
		return rc;
	return finishFrame(c);
}

}
