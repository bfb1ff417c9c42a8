// Drop-in minirender::Renderer on top of the CUDA pipeline.
//
// Host part of the reference's Renderer::render (reference src/Renderer.cpp:311-338): flatten
// the scene, derive the light vector, ambient term, projection kind and near plane, and form
// modelview / normal matrices per renderable with the same asl:: expressions. Everything after
// that (reference loops A..E, src/Renderer.cpp:344-380 and :163-309) runs on the GPU behind the
// C ABI of <minirender_b200.h>. No pixel is ever produced on the host.
#include <minirender/Renderer.h>
#include <minirender_b200.h>
#include "HostPool.h"
#include "HostInternal.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

using namespace asl;

namespace minirender {

// ---- projection builders (reference src/Renderer.cpp:40-83); `tan` is evaluated in double like
// the reference's unqualified call, then narrowed ----

Matrix4 projectionOrtho(float l, float r, float b, float t, float n, float f)
{
	const float w = r - l, h = t - b, d = f - n;
	return Matrix4(2 / w, 0, 0, -(r + l) / w,
	               0, 2 / h, 0, -(t + b) / h,
	               0, 0, -2 / d, -(f + n) / d,
	               0, 0, 0, 1);
}

Matrix4 projectionPerspective(float l, float r, float b, float t, float n, float f)
{
	const float w = r - l, h = t - b, d = f - n;
	return Matrix4(2 * n / w, 0, (r + l) / w, 0,
	               0, 2 * n / h, (t + b) / h, 0,
	               0, 0, -(f + n) / d, -2 * f * n / d,
	               0, 0, -1, 0);
}

Matrix4 projectionCV(const Matrix4& K, float w, float h, float n, float f)
{
	const float d = f - n;
	return Matrix4(K(0, 0) * 2 / w, K(0, 1) * 2 / w, -2 * K(0, 2) / w + 1, 0,
	               K(1, 0) * 2 / h, K(1, 1) * 2 / h, 2 * K(1, 2) / h - 1, 0,
	               0, 0, -(f + n) / d, -2 * f * n / d,
	               0, 0, -1, 0);
}

static inline float halfTan(float fov)
{
	return (float)::tan((double)(fov / 2));
}

Matrix4 projectionFrustum(float fov, float aspect, float n, float f)
{
	const float t = halfTan(fov);
	return projectionPerspective(-n * t * aspect, n * t * aspect, -n * t, n * t, n, f);
}

Matrix4 projectionFrustumH(float fov, float aspect, float n, float f)
{
	const float t = halfTan(fov);
	return projectionPerspective(-n * t, n * t, -n * t / aspect, n * t / aspect, n, f);
}

Matrix4 projectionOrtho(float fov, float aspect, float n, float f)
{
	return projectionOrtho(-fov * aspect / 2, fov * aspect / 2, -fov / 2, fov / 2, n, f);
}

// ---- implementation state ----

namespace {

// What identifies a mesh's geometry as mirrored in HBM: where its arrays are, how long they are, and a fingerprint of
// their contents. The reference re-reads every mesh on every render() (Renderer.cpp:341-380), so an edit in place
// (TriMesh::applyTransform, a deformed vertex array, a mesh freed and another allocated at the same address) shows
// in its next frame; here it has to be noticed. Hashing tens of megabytes every frame would cost more than the
// frame, so the fingerprint is taken from samples: arrays up to MR_FP_FULL bytes are hashed whole, longer ones at
// MR_FP_SAMPLES evenly spaced 16-byte words plus both ends. MINIRENDER_B200_GEOMETRY_CHECK=full hashes everything,
// =off compares addresses and lengths only; Renderer::invalidateGeometry() forces an upload in any mode.
struct MeshSig
{
	const void* p[6];
	int n[6];
	unsigned long long fp;
	bool operator==(const MeshSig& o) const { return memcmp(p, o.p, sizeof(p)) == 0 && memcmp(n, o.n, sizeof(n)) == 0 && fp == o.fp; }
};

enum { MR_FP_FULL = 4096, MR_FP_SAMPLES = 61, MR_FP_BUDGET = 1 << 20 };

static inline unsigned long long fpMix(unsigned long long h, unsigned long long v)
{
	h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
	return h * 0xff51afd7ed558ccdull;
}

// How much of an array a frame's fingerprint looks at. A scene of a few meshes gets MR_FP_FULL / MR_FP_SAMPLES; a scene
// of thousands shares MR_FP_BUDGET bytes per frame between its meshes (hashing every small array of a 10 000-mesh
// scene whole was measured at 27 ms per frame, twenty times the rest of render()'s host time).
struct FpPlan
{
	size_t full; // arrays up to this many bytes are hashed whole
	int samples; // longer ones: this many evenly spaced 16-byte words plus `edge` bytes at both ends
	size_t edge;
};

static FpPlan fingerprintPlan(size_t nMeshes)
{
	FpPlan p = { MR_FP_FULL, MR_FP_SAMPLES, 64 };
	const size_t perMesh = MR_FP_BUDGET / (nMeshes ? nMeshes : 1); // bytes per mesh (4 to 6 arrays)
	if (perMesh < 6 * (size_t)MR_FP_FULL)
	{
		p.full = std::max<size_t>(64, std::min<size_t>(MR_FP_FULL, perMesh / 6 / 16 * 16));
		p.samples = (int)std::max<size_t>(2, std::min<size_t>(MR_FP_SAMPLES, perMesh / 6 / 32));
		p.edge = 16;
	}
	return p;
}

static unsigned long long fingerprint(unsigned long long h, const void* ptr, size_t bytes, int mode, const FpPlan& plan)
{
	if (!ptr || !bytes || mode == 0)
		return h;
	const unsigned char* b = (const unsigned char*)ptr;
	unsigned long long w;
	if (mode == 2 || bytes <= plan.full || bytes < 2 * plan.edge + 32)
	{
		size_t i = 0;
		for (; i + 8 <= bytes; i += 8)
		{
			memcpy(&w, b + i, 8);
			h = fpMix(h, w);
		}
		for (; i < bytes; i++)
			h = fpMix(h, b[i]);
		return h;
	}
	for (size_t i = 0; i < plan.edge; i += 8) // both ends
	{
		memcpy(&w, b + i, 8);
		h = fpMix(h, w);
		memcpy(&w, b + bytes - plan.edge + i, 8);
		h = fpMix(h, w);
	}
	const size_t step = (bytes - 16) / (size_t)plan.samples;
	for (int k = 1; k < plan.samples; k++)
	{
		const size_t at = (size_t)k * step;
		memcpy(&w, b + at, 8);
		h = fpMix(h, w);
		memcpy(&w, b + at + 8, 8);
		h = fpMix(h, w);
	}
	return h;
}

// The lines fingerprint() will read of a long array, requested ahead of time: the arrays of a many-mesh scene are
// scattered over far more memory than the caches hold, and a frame's fingerprints are a few cache misses per array.
static inline void fingerprintPrefetch(const void* ptr, size_t bytes, int mode, const FpPlan& plan)
{
	if (!ptr || !bytes || mode == 0)
		return;
	const unsigned char* b = (const unsigned char*)ptr;
	if (mode == 2 || bytes <= plan.full || bytes < 2 * plan.edge + 32)
	{
		for (size_t i = 0; i < bytes && i < 512; i += 64)
			__builtin_prefetch(b + i);
		return;
	}
	__builtin_prefetch(b);
	__builtin_prefetch(b + bytes - plan.edge);
	const size_t step = (bytes - 16) / (size_t)plan.samples;
	for (int k = 1; k < plan.samples && k < 16; k++)
		__builtin_prefetch(b + (size_t)k * step);
}

static int geometryCheckMode()
{
	static int mode = -1;
	if (mode < 0)
	{
		const char* e = getenv("MINIRENDER_B200_GEOMETRY_CHECK");
		mode = (e && !strcmp(e, "off")) ? 0 : (e && !strcmp(e, "full")) ? 2 : 1;
	}
	return mode;
}

struct TexSig
{
	const void* p;
	int rows, cols;
	bool operator==(const TexSig& o) const { return p == o.p && rows == o.rows && cols == o.cols; }
};

void copy3x4(float* dst, const Matrix4& m)
{
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 4; j++)
			dst[4 * i + j] = m(i, j);
}

// Frames with at least this many renderables deal their host work to the pool (HostPool.h), in pieces of MR_POOL_CHUNK
enum { MR_POOL_MIN_ENTRIES = 1024, MR_POOL_CHUNK = 256 };

// One piece of a flatten dealt to the pool: a sub-tree (node, parent's world transform), or a mesh that the
// expansion of its parent already placed (emit = true: xform is its own world transform, children are separate items).
struct FlatItem
{
	SceneNode* node;
	Matrix4 xform;
	bool emit;
};

struct MatTex
{
	const void* p;
	int rows, cols;
};

}

struct Renderer::Impl
{
	mr_ctx* ctx;
	int device;
	int ctxW, ctxH;

	// flattened scene + descriptors of the current frame
	Array<Renderable> renderables;
	std::vector<mr_mesh_desc> meshes;
	std::vector<mr_texture_desc> textures;
	std::vector<mr_renderable> rlist;
	std::vector<mr_material> materials;
	mr_scene_desc sceneDesc;
	mr_frame frame;

	// Structure of the flattened list as describe() last saw it: the mesh / material of every entry and the
	// descriptor index it got, plus the distinct meshes / materials in descriptor order. While the list keeps
	// that structure (the usual case: only transforms and views change between frames) the descriptors are
	// refilled from these arrays without any look-up. Pointers are compared, never dereferenced, before
	// the fresh list has confirmed them.
	std::vector<const TriMesh*> cacheMeshOf;
	std::vector<const Material*> cacheMatOf;
	std::vector<int> cacheMeshId, cacheMatId;
	std::vector<const TriMesh*> cacheUniqMesh;
	std::vector<const Material*> cacheUniqMat;

	// Many-mesh scenes (HostPool.h): sub-trees flattened side by side, and each material's texture as first noted
	std::vector<FlatItem> flatItems;
	std::vector<std::vector<Renderable> > flatOut;
	std::vector<int> flatOffset;
	std::vector<MatTex> matTex;
	std::vector<MeshSig> sigScratch;
	int lastCount; // renderables of the previous frame: decides whether the next one is worth dealing to the pool

	// what is currently mirrored in HBM
	std::vector<MeshSig> upMeshes;
	std::vector<TexSig> upTextures;
	unsigned upStamp, upEpoch;
	bool uploaded;

	// frame constants snapshotted by render() (reference members _lightdir/_znear/_ambient)
	Vec3 lightdir;
	float znear, ambient;
	bool haveSnapshot;
	bool skipNearTest;   // paintTriangle(..., world = false) in progress
	bool buffersDefined; // the device buffers hold a frame or the clear values (the reference's ctor / setSize run clear())

	// lazily filled host mirrors of the device images
	Array2<Vec3> image, normals, points;
	Array2<float> depth;
	bool imageValid, depthValid, normalsValid;
	void* pinnedImage; // page-locked (mr_host_register) so that D2H runs at full PCIe rate
	void* pinnedDepth;

	Impl() : ctx(0), device(0), ctxW(0), ctxH(0), lastCount(0), upStamp(0), upEpoch(0), uploaded(false), znear(0), ambient(0.1f), haveSnapshot(false), skipNearTest(false), buffersDefined(false),
	         imageValid(false), depthValid(false), normalsValid(false), pinnedImage(0), pinnedDepth(0)
	{
		memset(&sceneDesc, 0, sizeof(sceneDesc));
		memset(&frame, 0, sizeof(frame));
		const char* env = getenv("MINIRENDER_B200_DEVICE");
		if (env)
			device = atoi(env);
	}
};

static void fail(mr_ctx* ctx, const char* what, int code)
{
	std::string msg = std::string("minirender_b200: ") + what + " failed (" + std::to_string(code) + ")";
	if (ctx && mr_last_error(ctx)[0])
		msg += std::string(": ") + mr_last_error(ctx);
	throw std::runtime_error(msg);
}

Renderer::Renderer() : _impl(new Impl), _w(800), _h(600)
{
	// defaults of the reference constructor (src/Renderer.cpp:85-98); the members the reference
	// leaves uninitialised (_saveNormals, _view) get defined values here
	_projection = projectionOrtho(-40, 40, -30, 30, 50, 120);
	_view = Matrix4::identity();
	_light = Vec3(-0.15f, 0.6f, 1).normalized();
	_defmaterial = new Material();
	_material = _defmaterial;
	_scene = NULL;
	_lighting = true;
	_texturing = true;
	_saveNormals = false;
	_lightIsPoint = false;
	_bgcolor = Vec3(0, 0, 0);
	_rowBegin = _rowEnd = 0;
	_geometryStamp = 0;
}

static void unpin(void*& p)
{
	if (p)
		mr_host_unregister(p);
	p = 0;
}

Renderer::~Renderer()
{
	unpin(_impl->pinnedImage);
	unpin(_impl->pinnedDepth);
	if (_impl->ctx)
		mr_destroy(_impl->ctx);
	delete _impl;
}

void Renderer::setDevice(int device)
{
	if (_impl->ctx && device != _impl->device)
	{
		mr_destroy(_impl->ctx);
		_impl->ctx = 0;
		_impl->uploaded = false;
	}
	_impl->device = device;
}

void Renderer::ensureContext()
{
	Impl& s = *_impl;
	if (!s.ctx)
	{
		int status = 0;
		s.ctx = mr_create(s.device, &status);
		if (!s.ctx)
			fail(0, "mr_create (this library has no CPU rasterizer; a CUDA device is required)", status);
		s.ctxW = s.ctxH = 0;
		s.uploaded = false;
		s.buffersDefined = false;
	}
	if (s.ctxW != _w || s.ctxH != _h)
	{
		s.buffersDefined = false;
		int rc = mr_set_size(s.ctx, _w, _h);
		if (rc)
			fail(s.ctx, "mr_set_size", rc);
		s.ctxW = _w;
		s.ctxH = _h;
	}
}

mr_ctx* Renderer::context()
{
	ensureContext();
	return _impl->ctx;
}

void Renderer::synchronize()
{
	if (_impl->ctx)
		mr_synchronize(_impl->ctx);
}

void Renderer::setSize(int w, int h)
{
	_w = w;
	_h = h;
	_impl->imageValid = _impl->depthValid = _impl->normalsValid = false;
	if (_impl->ctx) // the reference's setSize ends with clear()
		clear();
}

void Renderer::setScene(Shared<Scene> scene)
{
	_scene = scene;
}

// Geometry descriptor of one mesh: borrows the mesh's own arrays (asl::Vec3 / Vec2 are packed floats).
// false: the mesh cannot be drawn (TriMesh::normalsI shorter than TriMesh::indices).
static bool meshDescriptor(const TriMesh* mesh, mr_mesh_desc& d)
{
	memset(&d, 0, sizeof(d));
	if (mesh->normalsI.length() < mesh->indices.length())
		return false;
	d.positions = (const float*)mesh->vertices.ptr();
	d.n_positions = mesh->vertices.length();
	d.normals = (const float*)mesh->normals.ptr();
	d.n_normals = mesh->normals.length();
	d.idx_pos = mesh->indices.ptr();
	d.idx_nrm = mesh->normalsI.ptr();
	d.n_triangles = mesh->indices.length() / 3;
	// textured-capable only with both arrays present (reference Renderer.cpp:371)
	if (mesh->texcoords.length() > 0 && mesh->texcoordsI.length() >= mesh->indices.length() && mesh->texcoordsI.length() > 0)
	{
		d.texcoords = (const float*)mesh->texcoords.ptr();
		d.n_texcoords = mesh->texcoords.length();
		d.idx_uv = mesh->texcoordsI.ptr();
	}
	return true;
}

// Material descriptor without its texture index; what texture it has (if any) is noted in `tex`.
static void materialDescriptor(const Material* mat, mr_material& m, MatTex& tex)
{
	memset(&m, 0, sizeof(m));
	m.diffuse[0] = mat->diffuse.x; m.diffuse[1] = mat->diffuse.y; m.diffuse[2] = mat->diffuse.z;
	m.specular[0] = mat->specular.x; m.specular[1] = mat->specular.y; m.specular[2] = mat->specular.z;
	m.emissive[0] = mat->emissive.x; m.emissive[1] = mat->emissive.y; m.emissive[2] = mat->emissive.z;
	m.shininess = mat->shininess;
	m.texture = -1;
	tex.p = 0;
	tex.rows = tex.cols = 0;
	if (mat->texture.rows() > 0 && mat->texture.cols() > 0)
	{
		tex.p = mat->texture.data().ptr();
		tex.rows = mat->texture.rows();
		tex.cols = mat->texture.cols();
	}
}

// Builds descriptors for `list` (already flattened), in first-use order of the meshes and materials.
// Everything is re-read from the scene objects on every call (users mutate meshes, materials and
// transforms between frames); only the *structure* (which entry uses which mesh / material, and the index
// each distinct one got) is remembered from the previous call, so that a frame of 10 000 renderables does
// not pay 20 000 map look-ups (5.0 ms -> 1.3 ms on the cloud scene of configs[3]). Entries, meshes and materials
// are independent of each other: frames with >= MR_POOL_MIN_ENTRIES renderables deal them to the host pool
// (1.65 -> 0.4 ms for the same scene on 8 threads; the arithmetic per entry is the same code on any thread).
static void describe(Renderer::Impl& s, const Array<Renderable>& list, const Matrix4& view, Material* defmat)
{
	const int n = list.length();
	hostpool::Pool* const pool = n >= MR_POOL_MIN_ENTRIES ? hostpool::Pool::get() : 0;
	const Renderable* const entries = n ? &list[0] : 0;

	// ---- per entry: does it use the mesh / material it used last time; modelview and normal matrix ----
	const bool sized = (int)s.cacheMeshOf.size() == n && (int)s.cacheMatOf.size() == n && (int)s.cacheMeshId.size() == n && (int)s.cacheMatId.size() == n;
	std::atomic<int> differs(sized ? 0 : 1);
	s.rlist.resize((size_t)n);
	struct Entries
	{
		Renderer::Impl& s; const Renderable* list; const Matrix4& view; Material* defmat; int n; bool sized; std::atomic<int>& differs;
		void operator()(int c) const
		{
			const int i0 = c * MR_POOL_CHUNK, i1 = std::min(n, i0 + MR_POOL_CHUNK);
			bool same = sized;
			for (int i = i0; i < i1; i++)
			{
				const TriMesh* mesh = list[i].mesh;
				mr_renderable& r = s.rlist[(size_t)i];
				const Matrix4 modelview = view * list[i].transform;    // reference Renderer.cpp:337
				const Matrix4 normalmat = modelview.inverse().t();     // reference Renderer.cpp:338
				copy3x4(r.modelview, modelview);
				copy3x4(r.normalmat, normalmat);
				if (sized)
				{
					const Material* mat = mesh->material ? (Material*)mesh->material : defmat;
					same = same && mesh == s.cacheMeshOf[(size_t)i] && mat == s.cacheMatOf[(size_t)i];
					r.mesh = s.cacheMeshId[(size_t)i];
					r.material = s.cacheMatId[(size_t)i];
				}
			}
			if (!same)
				differs.store(1, std::memory_order_relaxed);
		}
	} perEntry = { s, entries, view, defmat, n, sized, differs };
	hostpool::parallelFor(pool, (n + MR_POOL_CHUNK - 1) / MR_POOL_CHUNK, perEntry);

	if (differs.load())
	{
		// the structure changed (or this is the first frame): number the distinct meshes and materials again
		std::map<const TriMesh*, int> meshIndex;
		std::map<const Material*, int> matIndex;
		s.cacheMeshOf.resize(n); s.cacheMatOf.resize(n); s.cacheMeshId.resize(n); s.cacheMatId.resize(n);
		s.cacheUniqMesh.clear(); s.cacheUniqMat.clear();
		for (int i = 0; i < n; i++)
		{
			TriMesh* mesh = list[i].mesh;
			const Material* mat = mesh->material ? (Material*)mesh->material : defmat;
			std::map<const TriMesh*, int>::iterator mi = meshIndex.find(mesh);
			if (mi == meshIndex.end())
			{
				mi = meshIndex.insert(std::make_pair(mesh, (int)s.cacheUniqMesh.size())).first;
				s.cacheUniqMesh.push_back(mesh);
			}
			std::map<const Material*, int>::iterator ti = matIndex.find(mat);
			if (ti == matIndex.end())
			{
				ti = matIndex.insert(std::make_pair(mat, (int)s.cacheUniqMat.size())).first;
				s.cacheUniqMat.push_back(mat);
			}
			s.cacheMeshOf[i] = mesh; s.cacheMatOf[i] = mat;
			s.cacheMeshId[i] = mi->second; s.cacheMatId[i] = ti->second;
			s.rlist[(size_t)i].mesh = mi->second;
			s.rlist[(size_t)i].material = ti->second;
		}
	}

	// ---- per distinct mesh / material: descriptors ----
	const int nMesh = (int)s.cacheUniqMesh.size(), nMat = (int)s.cacheUniqMat.size();
	s.meshes.resize((size_t)nMesh);
	s.materials.resize((size_t)nMat);
	s.matTex.resize((size_t)nMat);
	std::atomic<int> rejected(0);
	struct Descriptors
	{
		Renderer::Impl& s; int nMesh, nMat; std::atomic<int>& rejected;
		void operator()(int c) const
		{
			const int k0 = c * MR_POOL_CHUNK, k1 = k0 + MR_POOL_CHUNK;
			for (int k = k0; k < std::min(k1, nMesh); k++)
				if (!meshDescriptor(s.cacheUniqMesh[(size_t)k], s.meshes[(size_t)k]))
					rejected.store(1, std::memory_order_relaxed);
			for (int k = k0; k < std::min(k1, nMat); k++)
				materialDescriptor(s.cacheUniqMat[(size_t)k], s.materials[(size_t)k], s.matTex[(size_t)k]);
		}
	} perDistinct = { s, nMesh, nMat, rejected };
	hostpool::parallelFor(pool, (std::max(nMesh, nMat) + MR_POOL_CHUNK - 1) / MR_POOL_CHUNK, perDistinct);
	if (rejected.load())
	{
		// do not trust the cache after a rejected mesh
		s.cacheMeshOf.clear(); s.cacheMatOf.clear(); s.cacheMeshId.clear(); s.cacheMatId.clear();
		s.cacheUniqMesh.clear(); s.cacheUniqMat.clear();
		s.meshes.clear();
		throw std::runtime_error("minirender_b200: TriMesh::normalsI shorter than TriMesh::indices");
	}
	// textures in first-use order of the materials; the same texel array is listed once
	std::map<const void*, int> texIndex;
	s.textures.clear();
	for (int k = 0; k < nMat; k++)
	{
		const MatTex& t = s.matTex[(size_t)k];
		if (!t.p)
			continue;
		std::map<const void*, int>::iterator xi = texIndex.find(t.p);
		if (xi == texIndex.end())
		{
			mr_texture_desc d;
			d.texels = (const float*)t.p;
			d.rows = t.rows;
			d.cols = t.cols;
			xi = texIndex.insert(std::make_pair(t.p, (int)s.textures.size())).first;
			s.textures.push_back(d);
		}
		s.materials[(size_t)k].texture = xi->second;
	}

	s.sceneDesc.meshes = s.meshes.empty() ? 0 : &s.meshes[0];
	s.sceneDesc.n_meshes = (int)s.meshes.size();
	s.sceneDesc.textures = s.textures.empty() ? 0 : &s.textures[0];
	s.sceneDesc.n_textures = (int)s.textures.size();
	s.frame.renderables = s.rlist.empty() ? 0 : &s.rlist[0];
	s.frame.n_renderables = (int)s.rlist.size();
	s.frame.materials = s.materials.empty() ? 0 : &s.materials[0];
	s.frame.n_materials = (int)s.materials.size();
}

// ---- flatten (reference src/Scene.cpp:13-36) dealt to the host pool ----

// 2: exactly a TriMesh, 1: exactly one of the other node types of the API, 0: a type of the application's.
static inline int stockKind(const SceneNode* node)
{
	const std::type_info& t = typeid(*node);
	if (t == typeid(TriMesh))
		return 2;
	return (t == typeid(SceneNode) || t == typeid(Scene) || t == typeid(Shape)) ? 1 : 0;
}

// What SceneNode::collectShapes / TriMesh::collectShapes (host/Scene.cpp) do for a sub-tree of stock nodes, without the
// virtual calls; false as soon as a node of another type turns up (its own collectShapes has to run, on the caller's thread).
static bool flattenStock(SceneNode* node, const Matrix4& xform, std::vector<Renderable>& out)
{
	const int kind = stockKind(node);
	if (!kind)
		return false;
	const Matrix4 world = xform * node->transform;
	if (kind == 2)
		out.push_back(Renderable(static_cast<TriMesh*>(node), world));
	for (int i = 0; i < node->children.length(); i++)
		if (!flattenStock(node->children[i], world, out))
			return false;
	return true;
}

// The scene's renderables into s.renderables, in the order collectShapes gives them: the top levels of the graph are
// opened on this thread until there are enough sub-trees, each sub-tree is flattened into its own list by whichever
// thread takes it, and the lists are put behind each other. false: nothing usable was produced (a node type this
// library does not know, or too little to share out) and the caller runs collectShapes itself.
static bool flattenParallel(Renderer::Impl& s, Scene* scene, hostpool::Pool* pool)
{
	std::vector<FlatItem>& items = s.flatItems;
	std::vector<FlatItem> opened;
	items.clear();
	FlatItem root = { scene, Matrix4::identity(), false };
	items.push_back(root);
	const size_t want = (size_t)pool->width() * 4;
	for (int level = 0; level < 3 && items.size() < want; level++)
	{
		opened.clear();
		for (size_t k = 0; k < items.size(); k++)
		{
			const FlatItem& it = items[k];
			if (it.emit)
			{
				opened.push_back(it);
				continue;
			}
			const int kind = stockKind(it.node);
			if (!kind)
				return false;
			const Matrix4 world = it.xform * it.node->transform;
			if (kind == 2)
			{
				FlatItem e = { it.node, world, true };
				opened.push_back(e);
			}
			for (int i = 0; i < it.node->children.length(); i++)
			{
				FlatItem c = { it.node->children[i], world, false };
				opened.push_back(c);
			}
		}
		items.swap(opened);
	}
	const int nItems = (int)items.size();
	if (nItems < 2)
		return false;
	if ((int)s.flatOut.size() < nItems)
		s.flatOut.resize((size_t)nItems);
	std::atomic<int> foreign(0);
	struct Flatten
	{
		Renderer::Impl& s; std::atomic<int>& foreign;
		void operator()(int c) const
		{
			const FlatItem& it = s.flatItems[(size_t)c];
			std::vector<Renderable>& out = s.flatOut[(size_t)c];
			out.clear();
			if (it.emit)
				out.push_back(Renderable(static_cast<TriMesh*>(it.node), it.xform));
			else if (!foreign.load(std::memory_order_relaxed) && !flattenStock(it.node, it.xform, out))
				foreign.store(1, std::memory_order_relaxed);
		}
	} perItem = { s, foreign };
	hostpool::parallelFor(pool, nItems, perItem);
	if (foreign.load())
		return false;

	s.flatOffset.resize((size_t)nItems + 1);
	int total = 0;
	for (int k = 0; k < nItems; k++)
	{
		s.flatOffset[(size_t)k] = total;
		total += (int)s.flatOut[(size_t)k].size();
	}
	s.flatOffset[(size_t)nItems] = total;
	s.renderables.resize(total);
	if (total == 0)
		return true;
	struct Gather
	{
		Renderer::Impl& s; Renderable* dst; int total, nItems;
		void operator()(int c) const
		{
			const int i0 = c * MR_POOL_CHUNK, i1 = std::min(total, i0 + MR_POOL_CHUNK);
			// the item that holds entry i0: the last one that starts at or before it and is not empty
			int k = (int)(std::upper_bound(s.flatOffset.begin(), s.flatOffset.begin() + nItems, i0) - s.flatOffset.begin()) - 1;
			for (int i = i0; i < i1; k++)
			{
				const std::vector<Renderable>& src = s.flatOut[(size_t)k];
				const int first = i - s.flatOffset[(size_t)k];
				const int count = std::min((int)src.size() - first, i1 - i);
				for (int j = 0; j < count; j++)
					dst[i + j] = src[(size_t)(first + j)];
				i += count;
			}
		}
	} gather = { s, &s.renderables[0], total, nItems };
	hostpool::parallelFor(pool, (total + MR_POOL_CHUNK - 1) / MR_POOL_CHUNK, gather);
	return true;
}

static void fillFrameConstants(mr_frame& f, const Matrix4& P, const Vec3& lightdir, bool point, float ambient, float znear,
                               bool lighting, bool texturing, bool saveNormals, const Vec3& bg, int rowBegin, int rowEnd, int keep)
{
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++)
			f.projection[4 * i + j] = P(i, j);
	f.light[0] = lightdir.x; f.light[1] = lightdir.y; f.light[2] = lightdir.z;
	f.light_is_point = point;
	f.ambient = ambient;
	f.znear = znear;
	f.lighting = lighting;
	f.texturing = texturing;
	f.save_normals = saveNormals;
	f.background[0] = bg.x; f.background[1] = bg.y; f.background[2] = bg.z;
	f.row_begin = rowBegin;
	f.row_end = rowEnd;
	f.keep = keep;
}

void Renderer::prepare()
{
	Impl& s = *_impl;
	if (!_scene)
		throw std::runtime_error("minirender_b200: Renderer::render() without a scene");
	// (a frame as large as the last one was: sub-trees of stock nodes side by side on the host pool, HostPool.h)
	hostpool::Pool* const pool = s.lastCount >= MR_POOL_MIN_ENTRIES ? hostpool::Pool::get() : 0;
	if (!pool || !flattenParallel(s, _scene, pool))
	{
		s.renderables.clear();
		_scene->collectShapes(s.renderables, Matrix4::identity());
	}
	s.lastCount = s.renderables.length();

	// reference Renderer.cpp:317-326
	s.lightdir = _lightIsPoint ? _view * _light : _light.normalized();
	s.ambient = _scene->ambientLight;
	const bool persp = _projection(3, 3) == 0;
	float zn = persp ? _projection(2, 3) / (_projection(2, 2) - 1) : (_projection(2, 3) + 1) / _projection(2, 2);
	s.znear = -zn;
	s.haveSnapshot = true;

	describe(s, s.renderables, _view, _defmaterial);
	fillFrameConstants(s.frame, _projection, s.lightdir, _lightIsPoint, s.ambient, s.znear, _lighting, _texturing, _saveNormals,
	                   _bgcolor, _rowBegin, _rowEnd, 0);
}

const mr_scene_desc* Renderer::sceneDesc() const { return &_impl->sceneDesc; }
const mr_frame* Renderer::frameDesc() const { return &_impl->frame; }

// Uploads the geometry if what is mirrored in HBM does not match the descriptors.
static void syncGeometry(Renderer::Impl& s, unsigned stamp, bool force)
{
	const int mode = geometryCheckMode();
	const size_t nMeshes = s.meshes.size();
	const FpPlan plan = fingerprintPlan(nMeshes);
	std::vector<MeshSig>& ms = s.sigScratch;
	ms.resize(nMeshes);
	const bool comparable = !force && s.uploaded && s.upMeshes.size() == nMeshes;
	std::atomic<int> changed(comparable ? 0 : 1);
	struct Signatures
	{
		Renderer::Impl& s; std::vector<MeshSig>& ms; int mode; const FpPlan& plan; bool comparable; std::atomic<int>& changed;
		static void arrays(const mr_mesh_desc& d, const void* p[6], int n[6])
		{
			const void* pp[6] = { d.positions, d.normals, d.texcoords, d.idx_pos, d.idx_nrm, d.idx_uv };
			const int nn[6] = { d.n_positions, d.n_normals, d.n_texcoords, d.n_triangles, d.n_triangles, d.idx_uv ? d.n_triangles : 0 };
			memcpy(p, pp, sizeof(pp));
			memcpy(n, nn, sizeof(nn));
		}
		void operator()(int c) const
		{
			const size_t elem[6] = { 12, 12, 8, 12, 12, 12 };
			const size_t ahead = 3; // meshes whose sample lines are on their way while one is hashed
			const size_t i0 = (size_t)c * MR_POOL_CHUNK, i1 = std::min(ms.size(), i0 + MR_POOL_CHUNK);
			const void* p[6];
			int n[6];
			for (size_t i = i0; i < std::min(i1, i0 + ahead); i++)
			{
				arrays(s.meshes[i], p, n);
				for (int k = 0; k < 6; k++)
					fingerprintPrefetch(p[k], elem[k] * (size_t)n[k], mode, plan);
			}
			bool same = comparable;
			for (size_t i = i0; i < i1; i++)
			{
				if (i + ahead < i1)
				{
					arrays(s.meshes[i + ahead], p, n);
					for (int k = 0; k < 6; k++)
						fingerprintPrefetch(p[k], elem[k] * (size_t)n[k], mode, plan);
				}
				arrays(s.meshes[i], p, n);
				memset(&ms[i], 0, sizeof(MeshSig));
				memcpy(ms[i].p, p, sizeof(ms[i].p));
				memcpy(ms[i].n, n, sizeof(ms[i].n));
				unsigned long long h = 0x243f6a8885a308d3ull;
				for (int k = 0; k < 6; k++)
					h = fingerprint(h, p[k], elem[k] * (size_t)n[k], mode, plan);
				ms[i].fp = h;
				same = same && ms[i] == s.upMeshes[i];
			}
			if (!same)
				changed.store(1, std::memory_order_relaxed);
		}
	} perMesh = { s, ms, mode, plan, comparable, changed };
	hostpool::parallelFor(nMeshes >= (size_t)MR_POOL_MIN_ENTRIES ? hostpool::Pool::get() : 0, (int)((nMeshes + MR_POOL_CHUNK - 1) / MR_POOL_CHUNK), perMesh);
	std::vector<TexSig> ts(s.textures.size());
	for (size_t i = 0; i < s.textures.size(); i++)
	{
		ts[i].p = s.textures[i].texels;
		ts[i].rows = s.textures[i].rows;
		ts[i].cols = s.textures[i].cols;
	}
	const unsigned epoch = geometryEpoch().load(std::memory_order_acquire); // bumped by TriMesh::applyTransform()
	if (!changed.load() && stamp == s.upStamp && epoch == s.upEpoch && ts == s.upTextures)
		return;
	int rc = mr_upload_scene(s.ctx, &s.sceneDesc);
	if (rc)
		fail(s.ctx, "mr_upload_scene", rc);
	s.upMeshes.swap(ms);
	s.upTextures.swap(ts);
	s.upStamp = stamp;
	s.upEpoch = epoch;
	s.uploaded = true;
}

void Renderer::render()
{
	prepare();
	ensureContext();
	Impl& s = *_impl;
	const bool strip = _rowEnd > _rowBegin && (_rowBegin > 0 || _rowEnd < _h);
	if (strip && !s.buffersDefined)
		clear(); // a strip only writes its own rows: the others show the cleared buffers
	syncGeometry(s, _geometryStamp, false);
	int rc = mr_render(s.ctx, &s.frame);
	if (rc)
		fail(s.ctx, "mr_render", rc);
	s.buffersDefined = true;
	s.imageValid = s.depthValid = s.normalsValid = false;
}

void Renderer::clear()
{
	// reference Renderer.cpp:113-119: fill with background / 1e11f. On the device this is a
	// frame with no renderables (every tile is written once with the clear values).
	ensureContext();
	Impl& s = *_impl;
	mr_frame f;
	memset(&f, 0, sizeof(f));
	fillFrameConstants(f, _projection, Vec3(0, 0, 1), false, 0.1f, 0.f, _lighting, _texturing, _saveNormals, _bgcolor, 0, 0, 0);
	if (!s.uploaded)
	{
		mr_scene_desc empty;
		memset(&empty, 0, sizeof(empty));
		int rc = mr_upload_scene(s.ctx, &empty);
		if (rc)
			fail(s.ctx, "mr_upload_scene", rc);
		s.upMeshes.clear();
		s.upTextures.clear();
		s.uploaded = true;
		s.upStamp = _geometryStamp;
	}
	int rc = mr_render(s.ctx, &f);
	if (rc)
		fail(s.ctx, "mr_render(clear)", rc);
	s.buffersDefined = true;
	s.imageValid = s.depthValid = s.normalsValid = false;
}

// Immediate mode (reference Renderer.h:58-59): paints into the current buffers without
// clearing, using the frame constants of the last render() like the reference's members do.
void Renderer::paintMesh(TriMesh* mesh, const Matrix4& transform)
{
	ensureContext();
	Impl& s = *_impl;
	if (!s.buffersDefined)
		clear(); // painting into a fresh renderer starts from the cleared buffers (reference ctor / setSize)
	if (!s.haveSnapshot)
	{
		s.lightdir = _lightIsPoint ? _view * _light : _light.normalized();
		s.ambient = _scene ? _scene->ambientLight : 0.1f;
		const bool persp = _projection(3, 3) == 0;
		float zn = persp ? _projection(2, 3) / (_projection(2, 2) - 1) : (_projection(2, 3) + 1) / _projection(2, 2);
		s.znear = -zn;
	}
	Array<Renderable> one;
	one << Renderable(mesh, transform);
	describe(s, one, _view, _defmaterial);
	_material = mesh->material ? mesh->material : _defmaterial; // reference Renderer.cpp:336
	// (paintTriangle(..., world = false) is the reference's entry for its own clipper's output: no near test,
	// Renderer.cpp:169. A near plane no vertex can be in front of has the same effect on the device.)
	fillFrameConstants(s.frame, _projection, s.lightdir, _lightIsPoint, s.ambient, s.skipNearTest ? 3.0e38f : s.znear, _lighting, _texturing, _saveNormals,
	                   _bgcolor, _rowBegin, _rowEnd, 1);
	syncGeometry(s, _geometryStamp, true);
	int rc = mr_render(s.ctx, &s.frame);
	if (rc)
		fail(s.ctx, "mr_render(paintMesh)", rc);
	mr_synchronize(s.ctx); // the borrowed mesh may go away after this call
	s.uploaded = false;    // next render() re-mirrors the scene
	s.imageValid = s.depthValid = s.normalsValid = false;
}

void Renderer::paintTriangle(const Vertex& a, const Vertex& b, const Vertex& c, bool world)
{
	// Vertices are already in view space (reference Renderer.cpp:163-177): one-triangle mesh under an identity
	// modelview. With world == false the reference skips the near-plane test (its clipper re-enters that way).
	TriMesh tri;
	tri.vertices << a.position << b.position << c.position;
	tri.normals << a.normal << b.normal << c.normal;
	tri.texcoords << a.uv << b.uv << c.uv;
	tri.indices << 0 << 1 << 2;
	tri.normalsI = tri.indices;
	tri.texcoordsI = tri.indices;
	tri.material = _material;
	const Matrix4 savedView = _view;
	_view = Matrix4::identity();
	_impl->skipNearTest = !world;
	try
	{
		paintMesh(&tri, Matrix4::identity());
	}
	catch (...)
	{
		_view = savedView;
		_impl->skipNearTest = false;
		throw;
	}
	_impl->skipNearTest = false;
	_view = savedView;
}

Array2<Vec3> Renderer::getImage() const
{
	Impl& s = *_impl;
	if (!s.ctx || !s.buffersDefined) // a fresh renderer shows the cleared buffers, like the reference's (ctor / setSize end with clear())
		const_cast<Renderer*>(this)->clear();
	if (!s.imageValid)
	{
		if (s.image.rows() != _h || s.image.cols() != _w)
		{
			unpin(s.pinnedImage);
			s.image = Array2<Vec3>(_h, _w);
			if (mr_host_register(s.image.data().ptr(), sizeof(Vec3) * (size_t)_w * _h) == 0)
				s.pinnedImage = s.image.data().ptr();
		}
		int rc = mr_read_image(s.ctx, (float*)s.image.data().ptr());
		if (rc)
			fail(s.ctx, "mr_read_image", rc);
		s.imageValid = true;
	}
	return s.image;
}

Array2<float> Renderer::getDepth() const
{
	Impl& s = *_impl;
	if (!s.ctx || !s.buffersDefined) // a fresh renderer shows the cleared buffers, like the reference's (ctor / setSize end with clear())
		const_cast<Renderer*>(this)->clear();
	if (!s.depthValid)
	{
		if (s.depth.rows() != _h || s.depth.cols() != _w)
		{
			unpin(s.pinnedDepth);
			s.depth = Array2<float>(_h, _w);
			if (mr_host_register(s.depth.data().ptr(), sizeof(float) * (size_t)_w * _h) == 0)
				s.pinnedDepth = s.depth.data().ptr();
		}
		int rc = mr_read_depth(s.ctx, s.depth.data().ptr());
		if (rc)
			fail(s.ctx, "mr_read_depth", rc);
		s.depthValid = true;
	}
	return s.depth;
}

Array2<Vec3> Renderer::getNormalsImage() const
{
	Impl& s = *_impl;
	if (!s.ctx || !s.buffersDefined) // a fresh renderer shows the cleared buffers, like the reference's (ctor / setSize end with clear())
		const_cast<Renderer*>(this)->clear();
	if (!s.normalsValid)
	{
		if (s.normals.rows() != _h || s.normals.cols() != _w)
			s.normals = Array2<Vec3>(_h, _w);
		int rc = mr_read_normals(s.ctx, (float*)s.normals.data().ptr());
		if (rc)
			fail(s.ctx, "mr_read_normals", rc);
		s.normalsValid = true;
	}
	return s.normals;
}

Array2<Vec3> Renderer::getRangeImage()
{
	Impl& s = *_impl;
	if (!s.ctx)
		clear();
	if (s.points.rows() != _h || s.points.cols() != _w)
		s.points = Array2<Vec3>(_h, _w);
	float P[16];
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++)
			P[4 * i + j] = _projection(i, j);
	int rc = mr_read_range(s.ctx, P, s.znear, (float*)s.points.data().ptr());
	if (rc)
		fail(s.ctx, "mr_read_range", rc);
	return s.points;
}

Array<byte> Renderer::getImageRGB8() const
{
	Impl& s = *_impl;
	if (!s.ctx || !s.buffersDefined) // a fresh renderer shows the cleared buffers, like the reference's (ctor / setSize end with clear())
		const_cast<Renderer*>(this)->clear();
	Array<byte> out(_w * _h * 3);
	int rc = mr_read_rgb8(s.ctx, out.ptr());
	if (rc)
		fail(s.ctx, "mr_read_rgb8", rc);
	return out;
}

}
