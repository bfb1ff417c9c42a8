// minirender (B200 build) — procedural meshes with the reference's signatures
// (reference include/minirender/primitives.h:8-13). Host-side generators; their output
// arrays are identical to the reference's for the same arguments (tests/test_host.py).
#ifndef MINIRENDER_B200_PRIMITIVES_H
#define MINIRENDER_B200_PRIMITIVES_H

#include "Scene.h"

namespace minirender {

asl::Shared<TriMesh> createCube(float size = 1.0f);
asl::Shared<TriMesh> createCylinder(float radius, float height, int segments = 32, int heightSegments = 1, bool caps = true);
asl::Shared<TriMesh> createSphere(float radius, int latSegments = 16, int longSegments = 32);

}
#endif
