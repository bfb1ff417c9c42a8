// minirender (B200 build) — scene data model.
//
// Same type names, members and virtual interface as the reference's
// include/minirender/Scene.h:14-95, so user code that builds a scene for the reference
// recompiles against this header unchanged. The scene stays a plain host-side data
// structure; Renderer::render() flattens it (Scene.cpp, DFS pre-order like the reference's
// src/Scene.cpp:13-36) and mirrors the geometry into device memory.
#ifndef MINIRENDER_B200_SCENE_H
#define MINIRENDER_B200_SCENE_H

#include <asl/Array.h>
#include <asl/Array2.h>
#include <asl/Matrix4.h>
#include <asl/Pointer.h>
#include <asl/String.h>
#include <asl/Vec2.h>
#include <asl/Vec3.h>

namespace minirender {

struct TriMesh;

// A triangle corner as handed to Renderer::paintTriangle (reference Scene.h:14-24).
struct Vertex
{
	asl::Vec3 position;
	asl::Vec3 normal;
	asl::Vec2 uv;

	Vertex() {}
	Vertex(const asl::Vec3& p) : position(p), normal(0, 0, 1), uv(0, 0) {}
	Vertex(const asl::Vec3& p, const asl::Vec3& n) : position(p), normal(n), uv(0, 0) {}
	Vertex(const asl::Vec3& p, const asl::Vec3& n, const asl::Vec2& t) : position(p), normal(n), uv(t) {}
};

// Axis-aligned box grown by points/boxes (reference Scene.h:26-34).
struct BBox
{
	asl::Vec3 pmin, pmax;
	BBox()
	{
		pmin = asl::Vec3(1, 1, 1) * asl::infinity();
		pmax = -pmin;
	}
	BBox& operator+=(const asl::Vec3& p)
	{
		pmin = min(pmin, p);
		pmax = max(pmax, p);
		return *this;
	}
	BBox& operator+=(const BBox& b)
	{
		pmin = min(pmin, b.pmin);
		pmax = max(pmax, b.pmax);
		return *this;
	}
	asl::Vec3 size() const { return max(pmax - pmin, asl::Vec3::zeros()); }
	asl::Vec3 center() const { return (pmax + pmin) / 2; }
};

// Surface description (reference Scene.h:38-45, defaults src/Scene.cpp:65-71).
// `opacity` is carried for API compatibility; like the reference, the renderer ignores it.
struct Material
{
	asl::Vec3 diffuse, specular, emissive;
	float shininess, opacity;
	asl::Array2<asl::Vec3> texture; // float RGB, rows x cols; empty = untextured
	asl::String textureName;
	Material();
};

// One entry of the flattened scene: a mesh and its world transform (reference Scene.h:47-53).
struct Renderable
{
	TriMesh* mesh;
	asl::Matrix4 transform;
	Renderable() : mesh(0) {}
	Renderable(TriMesh* m, const asl::Matrix4& t) : mesh(m), transform(t) {}
};

struct SceneNode
{
	bool visible; // never read by the renderer (nor by the reference's)
	asl::Matrix4 transform;
	asl::Array<asl::Shared<SceneNode> > children;

	SceneNode();
	virtual ~SceneNode() {}
	// Appends (mesh, world) pairs in depth-first pre-order; world = xform * transform.
	virtual void collectShapes(asl::Array<Renderable>& list, const asl::Matrix4& xform);
	virtual BBox getBbox(const asl::Matrix4& xform = asl::Matrix4::identity()) const;
};

struct Shape : public SceneNode
{
	asl::Shared<Material> material;
	virtual ~Shape() {}
	virtual void applyTransform() {}
};

// Indexed triangle mesh with separate index streams for positions, normals and texcoords
// (reference Scene.h:73-87). A mesh is textured only if texcoords and texcoordsI are both
// non-empty (reference src/Renderer.cpp:371).
struct TriMesh : public Shape
{
	asl::Array<asl::Vec3> vertices;
	asl::Array<asl::Vec3> normals;
	asl::Array<asl::Vec2> texcoords;
	asl::Array<int> indices;
	asl::Array<int> normalsI;
	asl::Array<int> texcoordsI;

	TriMesh();
	virtual void collectShapes(asl::Array<Renderable>& list, const asl::Matrix4& xform);
	virtual BBox getBbox(const asl::Matrix4& xform = asl::Matrix4::identity()) const;
	virtual void applyTransform();
};

struct Scene : public SceneNode
{
	float ambientLight;
	asl::Vec3 light; // unused by the renderer (the light is set on the Renderer)
	Scene();
	void add(const asl::Shared<SceneNode>& node) { children << node; }
};

}

#endif
