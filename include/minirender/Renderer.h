// minirender (B200 build) — drop-in Renderer.
//
// Public API identical to the reference's class (reference include/minirender/Renderer.h:9-64):
// same free projection builders, same setters, render(), and the same four image getters.
// The difference is where the work happens: render() flattens the scene on the host exactly
// like the reference (src/Renderer.cpp:313-338), then hands the frame to the CUDA pipeline
// through the C ABI in <minirender_b200.h>. Images live in HBM and are copied to the host
// lazily, when a getter is called. There is no CPU rasterizer in this library: constructing a
// Renderer without a CUDA device throws std::runtime_error on first use.
#ifndef MINIRENDER_B200_RENDERER_H
#define MINIRENDER_B200_RENDERER_H

#include "Scene.h"
#include <asl/Array2.h>

struct mr_ctx;
struct mr_frame;
struct mr_scene_desc;

namespace minirender {

asl::Matrix4 projectionOrtho(float l, float r, float b, float t, float n, float f);
asl::Matrix4 projectionOrtho(float fov, float aspect, float n, float f);
asl::Matrix4 projectionPerspective(float l, float r, float b, float t, float n, float f);
asl::Matrix4 projectionFrustum(float fov, float aspect, float n, float f);
asl::Matrix4 projectionFrustumH(float fov, float aspect, float n, float f);
asl::Matrix4 projectionCV(const asl::Matrix4& K, float w, float h, float n, float f);

class Renderer
{
public:
	Renderer();
	~Renderer();

	void setSize(int w, int h);
	float aspect() const { return (float)_w / _h; }
	void setScene(asl::Shared<Scene> scene);
	void setProjection(const asl::Matrix4& m) { _projection = m; }
	void setView(const asl::Matrix4& m) { _view = m; }
	void setLight(const asl::Vec3& v, bool point = false) { _light = v; _lightIsPoint = point; }
	void setMaterial(asl::Shared<Material> material) { _material = material; }
	void setLighting(bool on) { _lighting = on; }
	void setTexturing(bool on) { _texturing = on; }
	void setSaveNormals(bool on) { _saveNormals = on; }
	void setBackground(const asl::Vec3& color) { _bgcolor = color; }
	void clear();
	void render();
	void paintMesh(TriMesh* mesh, const asl::Matrix4& transform = asl::Matrix4::identity());
	void paintTriangle(const Vertex& a, const Vertex& b, const Vertex& c, bool world = true);
	asl::Array2<float>     getDepth() const;
	asl::Array2<asl::Vec3> getImage() const;
	asl::Array2<asl::Vec3> getRangeImage();
	asl::Array2<asl::Vec3> getNormalsImage() const;

	// ---- extensions (not in the reference) ----
	// Select the GPU before first use (default 0, or $MINIRENDER_B200_DEVICE).
	void setDevice(int device);
	// Geometry arrays are mirrored in HBM and re-uploaded when their storage or length changes.
	// Call this after editing vertex data *in place* (same storage, same length).
	void invalidateGeometry() { _geometryStamp++; }
	// Render only rows [begin,end) (strip sharding); 0,0 restores the whole image.
	void setRowRange(int begin, int end) { _rowBegin = begin; _rowEnd = end; }
	// 8-bit image quantised on the device with savePPM's rule (3 bytes/pixel over PCIe).
	asl::Array<asl::byte> getImageRGB8() const;
	// Host-side part of render() only: flatten + per-frame constants, no GPU work.
	// The descriptors stay valid until the next prepare()/render() on this object.
	void prepare();
	const mr_scene_desc* sceneDesc() const;
	const mr_frame* frameDesc() const;
	mr_ctx* context();
	void synchronize();

	struct Impl; // host-side state behind the API (defined in Renderer.cpp)

private:
	Renderer(const Renderer&);
	Renderer& operator=(const Renderer&);
	Impl* _impl;
	int _w, _h;
	asl::Matrix4 _view, _projection;
	asl::Vec3 _light, _bgcolor;
	asl::Shared<Scene> _scene;
	asl::Shared<Material> _material, _defmaterial;
	bool _lighting, _texturing, _lightIsPoint, _saveNormals;
	int _rowBegin, _rowEnd;
	unsigned _geometryStamp;
	void ensureContext();
	void flushImmediate();
};

}
#endif
