// minirender (B200 build) — file formats either side of the render path: PPM images / textures
// (reference include/minirender/io.h:18-20, src/io.cpp:337-415) and the mesh loaders / exporters
// (reference include/minirender/io.h:10-16,22; src/io.cpp:16-335, src/x3d.cpp). Same names and
// signatures as the reference's io.h.
#ifndef MINIRENDER_B200_IO_H
#define MINIRENDER_B200_IO_H

#include "Scene.h"

namespace minirender {

// Loads .stl / .obj / .x3d by extension into a scene subtree (reference src/io.cpp:113-132).
// Unknown extensions give an empty node; a file that cannot be read gives a null pointer.
asl::Shared<SceneNode> loadMesh(const asl::String& filename);

// ASCII or binary STL (binary iff the file size matches the facet count, src/io.cpp:134-156):
// three fresh vertices and one normal per facet.
asl::Shared<TriMesh> loadSTL(const asl::String& filename);

// Binary STL of the mesh's triangles with mesh->transform applied (src/io.cpp:158-185).
void saveSTL(asl::Shared<TriMesh> mesh, const asl::String& name);

// Wavefront OBJ + MTL (src/io.cpp:189-335): one TriMesh per material sharing the file's arrays,
// polygons fan-triangulated, v flipped, Kd/Ks/Ke/Ns/d/map_Kd (PPM) materials, flat normals when
// the file has none.
asl::Shared<SceneNode> loadOBJ(const asl::String& filename);

// X3D (src/x3d.cpp:35-203): Transform / Group hierarchy, Shape with Material / ImageTexture,
// IndexedFaceSet / IndexedTriangleSet, DEF / USE, Inline.
asl::Shared<SceneNode> loadX3D(const asl::String& filename);

// Fan triangulation of -1-terminated polygons (src/x3d.cpp:17-33).
asl::Array<int> triangulateIndices(const asl::Array<int>& indices);

// Range image (Renderer::getRangeImage) as "x y z" text lines, points in front of the camera only,
// transformed by m (src/io.cpp:417-431).
void saveXYZ(const asl::Array2<asl::Vec3>& points, const asl::String& filename, const asl::Matrix4& m = asl::Matrix4::identity());

// Binary P6 writer; each channel is (byte)clamp(v*255, 0, 255), i.e. truncation
// (reference src/io.cpp:358-361). "--" writes to stdout.
void savePPM(const asl::Array2<asl::Vec3>& image, const asl::String& filename);

// Binary P6 reader: '#' comments allowed in the header, maxval ignored, texel = rgb/255
// (reference src/io.cpp:367-415). Returns an empty image on failure.
asl::Array2<asl::Vec3> loadPPM(const asl::String& filename);

// The 8-bit quantiser alone (what parity on RGB is judged in).
void quantizeRGB8(const asl::Array2<asl::Vec3>& image, asl::byte* rgb8);

}
#endif
