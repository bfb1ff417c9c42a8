// Scene graph: constructors, flattening and bounding boxes.
// Behaviour follows the reference's src/Scene.cpp (flatten :13-36, bbox :22-47,
// applyTransform :53-63, defaults :65-76); the flatten order (depth-first, parent before
// children, a TriMesh emits itself before its own children) is what defines the submission
// order the rasterizer's depth ties are resolved in.
#include <minirender/Scene.h>
#include "HostInternal.h"

using asl::Array;
using asl::Matrix4;
using asl::Vec3;

namespace minirender {

std::atomic<unsigned>& geometryEpoch()
{
	static std::atomic<unsigned> epoch(0);
	return epoch;
}

SceneNode::SceneNode() : visible(true), transform(Matrix4::identity()) {}

void SceneNode::collectShapes(Array<Renderable>& list, const Matrix4& xform)
{
	const Matrix4 world = xform * transform;
	for (int i = 0; i < children.length(); i++)
		children[i]->collectShapes(list, world);
}

BBox SceneNode::getBbox(const Matrix4& xform) const
{
	BBox box;
	for (int i = 0; i < children.length(); i++)
		box += children[i]->getBbox(xform * transform);
	return box;
}

TriMesh::TriMesh()
{
	material = NULL;
}

void TriMesh::collectShapes(Array<Renderable>& list, const Matrix4& xform)
{
	const Matrix4 world = xform * transform;
	list << Renderable(this, world);
	for (int i = 0; i < children.length(); i++)
		children[i]->collectShapes(list, world);
}

BBox TriMesh::getBbox(const Matrix4& xform) const
{
	BBox box;
	for (int i = 0; i < vertices.length(); i++)
		box += xform * transform * vertices[i];
	for (int i = 0; i < children.length(); i++)
		box += children[i]->getBbox(xform * transform);
	return box;
}

void TriMesh::applyTransform()
{
	for (int i = 0; i < vertices.length(); i++)
		vertices[i] = transform * vertices[i];
	const Matrix4 nm = transform.inverse().transposed();
	for (int i = 0; i < normals.length(); i++)
		normals[i] = (nm % normals[i]).normalized();
	transform = Matrix4::identity();
	geometryEpoch().fetch_add(1u, std::memory_order_release); // the arrays changed in place (host/Renderer.cpp: syncGeometry)
}

Material::Material()
	: diffuse(0.7f, 0.7f, 0.9f), specular(0.8f, 0.8f, 0.8f), emissive(0, 0, 0), shininess(12.0f), opacity(1.0f)
{
}

Scene::Scene() : ambientLight(0.1f) {}

}
