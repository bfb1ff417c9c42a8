// Procedural meshes. Output arrays (vertex order, index order, float values) are the same as
// the reference generators' for the same arguments (reference src/primitives.cpp: cube :7-48,
// cylinder :50-148, sphere :150-230) — tests/test_host.py checks that bit for bit — because
// benchmark scenes (1M / 2M-triangle spheres) are specified through these calls.
#include <minirender/primitives.h>
#include <cmath>

using asl::Shared;
using asl::Vec2;
using asl::Vec3;

namespace minirender {

namespace {

struct MeshBuilder
{
	TriMesh* m;
	explicit MeshBuilder(TriMesh* mesh) : m(mesh) {}
	int vertex(const Vec3& p, const Vec3& n, const Vec2& t)
	{
		m->vertices << p;
		m->normals << n;
		m->texcoords << t;
		return m->vertices.length() - 1;
	}
	// positions and normals share one index stream in all primitives
	void tri(int a, int b, int c)
	{
		m->indices << a << b << c;
		m->normalsI << a << b << c;
	}
};

}

Shared<TriMesh> createCube(float size)
{
	TriMesh* mesh = new TriMesh();
	MeshBuilder mb(mesh);
	const float h = size * 0.5f;
	// Six faces in the order +X -X +Y -Y +Z -Z; four corners each (as sign patterns), wound so
	// that triangles (0,2,1) and (0,3,2) face outwards. Vertices are duplicated per face for
	// flat normals: 24 vertices, 12 triangles.
	static const signed char corner[6][4][3] = {
		{ { 1, -1, -1 }, { 1, -1, 1 }, { 1, 1, 1 }, { 1, 1, -1 } },
		{ { -1, -1, 1 }, { -1, -1, -1 }, { -1, 1, -1 }, { -1, 1, 1 } },
		{ { -1, 1, -1 }, { 1, 1, -1 }, { 1, 1, 1 }, { -1, 1, 1 } },
		{ { -1, -1, 1 }, { 1, -1, 1 }, { 1, -1, -1 }, { -1, -1, -1 } },
		{ { 1, -1, 1 }, { -1, -1, 1 }, { -1, 1, 1 }, { 1, 1, 1 } },
		{ { -1, -1, -1 }, { 1, -1, -1 }, { 1, 1, -1 }, { -1, 1, -1 } },
	};
	static const signed char facing[6][3] = { { 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } };
	static const float uvs[4][2] = { { 0, 0 }, { 1, 0 }, { 1, 1 }, { 0, 1 } };
	for (int f = 0; f < 6; f++)
	{
		const Vec3 n((float)facing[f][0], (float)facing[f][1], (float)facing[f][2]);
		int first = 0;
		for (int c = 0; c < 4; c++)
		{
			const Vec3 p(corner[f][c][0] > 0 ? h : -h, corner[f][c][1] > 0 ? h : -h, corner[f][c][2] > 0 ? h : -h);
			const int id = mb.vertex(p, n, Vec2(uvs[c][0], uvs[c][1]));
			if (c == 0)
				first = id;
		}
		mb.tri(first, first + 2, first + 1);
		mb.tri(first, first + 3, first + 2);
	}
	// note: no texcoordsI, so a cube is never textured by paintMesh (reference Renderer.cpp:371)
	mesh->material = new Material();
	return mesh;
}

Shared<TriMesh> createCylinder(float radius, float height, int segments, int heightSegments, bool caps)
{
	TriMesh* mesh = new TriMesh();
	MeshBuilder mb(mesh);
	if (segments < 3)
		segments = 3;
	if (heightSegments < 1)
		heightSegments = 1;
	const float halfH = height * 0.5f;
	const float twoPi = 2.0f * asl::PIf;

	// side wall: (heightSegments+1) rings of `segments` vertices, axis along z
	for (int ring = 0; ring <= heightSegments; ring++)
	{
		const float v = (float)ring / (float)heightSegments;
		const float z = -halfH + v * height;
		for (int s = 0; s < segments; s++)
		{
			const float u = (float)s / (float)segments;
			const float angle = u * twoPi;
			const float x = radius * std::cos(angle);
			const float y = radius * std::sin(angle);
			mb.vertex(Vec3(x, y, z), Vec3(x, y, 0).normalized(), Vec2(u, v));
		}
	}
	for (int ring = 0; ring < heightSegments; ring++)
		for (int s = 0; s < segments; s++)
		{
			const int s1 = (s + 1) % segments;
			const int lo0 = ring * segments + s, lo1 = ring * segments + s1;
			const int hi0 = lo0 + segments, hi1 = lo1 + segments;
			mb.tri(lo0, lo1, hi0);
			mb.tri(lo1, hi1, hi0);
		}

	if (caps)
	{
		// each cap: a copy of the rim ring with the cap normal, then the centre vertex, then a fan
		for (int top = 0; top < 2; top++)
		{
			const float nz = top ? 1.0f : -1.0f;
			const int rim = top ? heightSegments * segments : 0;
			const int start = mesh->vertices.length();
			for (int s = 0; s < segments; s++)
			{
				const Vec3 p = mesh->vertices[rim + s];
				mb.vertex(p, Vec3(0, 0, nz), Vec2((p.x / radius + 1.0f) * 0.5f, (p.y / radius + 1.0f) * 0.5f));
			}
			const int centre = mb.vertex(Vec3(0, 0, top ? halfH : -halfH), Vec3(0, 0, nz), Vec2(0.5f, 0.5f));
			for (int s = 0; s < segments; s++)
			{
				const int a = start + s, b = start + (s + 1) % segments;
				if (top)
					mb.tri(centre, a, b);
				else
					mb.tri(centre, b, a);
			}
		}
	}
	mesh->material = new Material();
	return mesh;
}

Shared<TriMesh> createSphere(float radius, int latSegments, int longSegments)
{
	TriMesh* mesh = new TriMesh();
	MeshBuilder mb(mesh);
	if (latSegments < 2)
		latSegments = 2;
	if (longSegments < 3)
		longSegments = 3;
	const int rings = latSegments - 1;
	mesh->vertices.reserve(rings * longSegments + 2);
	mesh->normals.reserve(rings * longSegments + 2);
	mesh->texcoords.reserve(rings * longSegments + 2);
	mesh->indices.reserve(6 * longSegments * rings);
	mesh->normalsI.reserve(6 * longSegments * rings);

	// y-up UV sphere: north pole, (latSegments-1) rings, south pole
	const int north = mb.vertex(Vec3(0.0f, radius, 0.0f), Vec3(0.0f, 1.0f, 0.0f), Vec2(0.5f, 0.0f));
	const int ring0 = north + 1;
	for (int lat = 1; lat < latSegments; lat++)
	{
		const float v = (float)lat / (float)latSegments;
		const float phi = v * asl::PIf;
		const float sinPhi = std::sin(phi), cosPhi = std::cos(phi);
		for (int lon = 0; lon < longSegments; lon++)
		{
			const float u = (float)lon / (float)longSegments;
			const float theta = u * 2.0f * asl::PIf;
			const float x = radius * sinPhi * std::cos(theta);
			const float y = radius * cosPhi;
			const float z = radius * sinPhi * std::sin(theta);
			mb.vertex(Vec3(x, y, z), Vec3(x, y, z).normalized(), Vec2(u, 1.0f - v));
		}
	}
	const int south = mb.vertex(Vec3(0.0f, -radius, 0.0f), Vec3(0.0f, -1.0f, 0.0f), Vec2(0.5f, 1.0f));

	for (int lon = 0; lon < longSegments; lon++) // north fan
		mb.tri(north, ring0 + lon, ring0 + (lon + 1) % longSegments);
	for (int r = 0; r + 1 < rings; r++) // quad bands, two triangles per quad
	{
		const int upper = ring0 + r * longSegments, lower = upper + longSegments;
		for (int lon = 0; lon < longSegments; lon++)
		{
			const int next = (lon + 1) % longSegments;
			mb.tri(upper + lon, upper + next, lower + lon);
			mb.tri(upper + next, lower + next, lower + lon);
		}
	}
	const int last = ring0 + (rings - 1) * longSegments;
	for (int lon = 0; lon < longSegments; lon++) // south fan
		mb.tri(south, last + lon, last + (lon + 1) % longSegments);

	mesh->material = new Material();
	return mesh;
}

}
