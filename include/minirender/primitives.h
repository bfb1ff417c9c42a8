// minirender (B200 build) — procedural meshes with the reference's signatures
// (reference include/minirender/primitives.h:8-13). Host-side generators: benchmark scenes are
// specified through them (createSphere(100, 501, 1000) is the 1,000,000-triangle mesh of
// BASELINE.json configs[1]), so their output arrays are identical to the reference generators' for
// the same arguments — vertex order, index order and float values (tests/test_host.py).
//
//   createCube      24 vertices / 12 triangles, flat normals, per-face uv in [0,1]^2 but no
//                   texcoordsI (so a cube is never textured)
//   createCylinder  axis along z, (heightSegments+1) rings of `segments` vertices, optional caps
//   createSphere    y-up UV sphere: (lat-1)*lon + 2 vertices, 2*lon*(lat-1) triangles
#ifndef MINIRENDER_B200_PRIMITIVES_H
#define MINIRENDER_B200_PRIMITIVES_H

#include "Scene.h"

namespace minirender {

asl::Shared<TriMesh> createCube(float size = 1.0f);
asl::Shared<TriMesh> createCylinder(float radius, float height, int segments = 32, int heightSegments = 1, bool caps = true);
asl::Shared<TriMesh> createSphere(float radius, int latSegments = 16, int longSegments = 32);

}
#endif
