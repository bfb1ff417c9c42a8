// minirender (B200 build) — scene data model, umbrella header.
//
// Same type names, members and virtual interface as the reference's single header
// include/minirender/Scene.h:14-95 (Vertex, BBox, Material, Renderable, SceneNode, Shape, TriMesh,
// Scene), so user code that builds a scene for the reference recompiles unchanged. The types
// live in Vertex.h, Material.h and SceneNode.h; this header adds the root node.
#ifndef MINIRENDER_B200_SCENE_H
#define MINIRENDER_B200_SCENE_H

#include "SceneNode.h"

namespace minirender {

// Root of the graph. `ambientLight` (default 0.1) is added to the diffuse term of every pixel;
// `light` is unused by the renderer (the light is set on the Renderer), kept for API compatibility.
struct Scene : public SceneNode
{
	float ambientLight;
	asl::Vec3 light;

	Scene();
	void add(const asl::Shared<SceneNode>& node) { children << node; }
};

}
#endif
