// minirender (B200 build) — small value types of the scene model: Vertex and BBox
// (API of reference include/minirender/Scene.h:14-34).
#ifndef MINIRENDER_B200_VERTEX_H
#define MINIRENDER_B200_VERTEX_H

#include <asl/Vec2.h>
#include <asl/Vec3.h>

namespace minirender {

// One triangle corner as handed to Renderer::paintTriangle: position and normal in the space the
// call expects (view space for paintTriangle), plus a texture coordinate.
// Defaults when parts are omitted: normal (0,0,1), uv (0,0).
struct Vertex
{
	asl::Vec3 position;
	asl::Vec3 normal;
	asl::Vec2 uv;

	Vertex() {}
	Vertex(const asl::Vec3& p, const asl::Vec3& n = asl::Vec3(0, 0, 1), const asl::Vec2& t = asl::Vec2(0, 0))
		: position(p), normal(n), uv(t)
	{
	}
};

// Axis-aligned bounding box that grows by points or other boxes. A fresh box is empty
// (pmin = +inf, pmax = -inf); size() of an empty box is zero.
struct BBox
{
	asl::Vec3 pmin, pmax;

	BBox() : pmin(asl::infinity(), asl::infinity(), asl::infinity()), pmax(-asl::infinity(), -asl::infinity(), -asl::infinity()) {}

	BBox& operator+=(const asl::Vec3& point)
	{
		pmin = min(pmin, point);
		pmax = max(pmax, point);
		return *this;
	}
	BBox& operator+=(const BBox& other)
	{
		pmin = min(pmin, other.pmin);
		pmax = max(pmax, other.pmax);
		return *this;
	}
	asl::Vec3 size() const { return max(pmax - pmin, asl::Vec3::zeros()); }
	asl::Vec3 center() const { return (pmax + pmin) / 2; }
};

}
#endif
