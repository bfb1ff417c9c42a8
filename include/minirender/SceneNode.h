// minirender (B200 build) — scene graph nodes (API of reference include/minirender/Scene.h:47-87).
// The graph stays a plain host-side structure. Renderer::render() flattens it depth first, parent
// before children, into (mesh, world transform) pairs; that order is the submission order in
// which equal-depth fragments are resolved.
#ifndef MINIRENDER_B200_SCENENODE_H
#define MINIRENDER_B200_SCENENODE_H

#include "Material.h"
#include "Vertex.h"
#include <asl/Array.h>
#include <asl/Matrix4.h>
#include <asl/Pointer.h>

namespace minirender {

struct TriMesh;

// One entry of the flattened scene. `mesh` is a raw pointer into the graph: valid only while the
// graph is alive (the renderer uses it during render() only).
struct Renderable
{
	TriMesh* mesh;
	asl::Matrix4 transform; // world transform: product of the node transforms from the root down

	Renderable() : mesh(0) {}
	Renderable(TriMesh* m, const asl::Matrix4& world) : mesh(m), transform(world) {}
};

// Interior node: a transform and children.
struct SceneNode
{
	bool visible; // kept for API compatibility: neither this renderer nor the reference reads it
	asl::Matrix4 transform;
	asl::Array<asl::Shared<SceneNode> > children;

	SceneNode();
	virtual ~SceneNode() {}
	// Appends this subtree's (mesh, world) pairs to `list`; world = xform * transform.
	virtual void collectShapes(asl::Array<Renderable>& list, const asl::Matrix4& xform);
	virtual BBox getBbox(const asl::Matrix4& xform = asl::Matrix4::identity()) const;
};

// A node that can carry a material.
struct Shape : public SceneNode
{
	asl::Shared<Material> material; // null: the renderer's default material
	virtual ~Shape() {}
	virtual void applyTransform() {}
};

// Indexed triangle mesh with separate index streams for positions, normals and texcoords:
// corner c of triangle t uses vertices[indices[3t+c]], normals[normalsI[3t+c]] and — only if both
// texcoords and texcoordsI are non-empty — texcoords[texcoordsI[3t+c]]. normalsI must be as long as
// indices. A mesh is itself a node: it emits itself first, then its children.
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
	// Bakes `transform` into the vertex / normal arrays and resets it to identity.
	virtual void applyTransform();
};

}
#endif
