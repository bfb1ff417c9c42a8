// bench_scene: the reference's own benchmark program (reference samples/bench.cpp) on the public
// minirender C++ API only: 20 revolved sinc-profile objects of 200 x 200 vertices (ring 0 is NaN by
// construction, as upstream: sin(0)/0), random placement, one directional light, 1920 x 1080,
// frustum 35 degrees, a turntable of n frames; prints the total and per-frame time like the original.
// Differences from the upstream file: plain argv parsing instead of asl::CmdArgs, and a seeded
// std::mt19937 instead of asl::Random (not available offline) — so the object placement is
// deterministic but not upstream's. Builds unchanged against this repo's headers +
// libminirender_b200.so (frames rendered on the GPU) and against the reference's headers + sources.
//
//   bench_scene [-n frames] [-w width] [-h height] [-d distance] [-rz deg/s] [-rx deg/s]
//               [-yaw deg] [-tilt deg] [-tex] [-dark] [-save] [-objects k] [-m rings] [-seg segments]
#include <minirender/Renderer.h>
#include <minirender/Scene.h>
#ifndef BENCH_SCENE_NO_IO // (the reference's io.cpp needs asl::File, which the offline stand-in does not provide)
#include <minirender/io.h>
#endif
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

using namespace asl;
using namespace minirender;

static std::mt19937 g_rng(1234);
static float rnd(float a, float b) { return a + (b - a) * (float)(g_rng() >> 8) * (1.0f / 16777216.0f); }
static float rnd(float b) { return rnd(0.0f, b); }

static double nowSeconds()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// profile of the revolved body (reference samples/bench.cpp:10-13)
static float rf(float z, float s)
{
	return 30 + 40 * sin(4 * z * (float)PI / s) / (4 * z * (float)PI / s);
}

// reference samples/bench.cpp:15-62
static Shared<TriMesh> createObject(int m, int n, bool usetex)
{
	Shared<TriMesh> mesh = new TriMesh();
	const float da = 2 * (float)PI / n;
	const float dz = 70.f / m;
	for (int i = 0; i < m; i++)
	{
		const float z = i * dz;
		const float r = rf(z, 100);
		const Vec2 nor = Vec2(1, -(rf(z + 0.001f, 100) - rf(z, 100)) / 0.001f).normalized();
		for (int j = 0; j < n; j++)
		{
			mesh->vertices << Vec3(r * cos(j * da), r * sin(j * da), z);
			mesh->normals << Vec3(nor.x * cos(j * da), nor.x * sin(j * da), nor.y).normalized();
			if (usetex)
				mesh->texcoords << Vec2((float)j / n, (float)i / m);
			if (j > 0 && i > 0)
				mesh->indices << (n * (i - 1) + j - 1) << (n * (i - 1) + j) << (n * i + j) << (n * (i - 1) + j - 1) << (n * i + j)
				              << (n * i + j - 1);
		}
	}
	mesh->normalsI = mesh->indices;
	if (usetex)
		mesh->texcoordsI = mesh->indices;
	mesh->material = new Material();
	mesh->material->shininess = 15;
	if (usetex)
	{
		const Vec3 color(rnd(1.f), rnd(1.f), rnd(1.f));
		Array2<Vec3> tex(256, 256);
		for (int i = 0; i < tex.rows(); i++)
			for (int j = 0; j < tex.cols(); j++)
				tex(i, j) = color * (0.75f + 0.25f * (cos(i * 40 / 256.f) * sin(j * 40 / 256.f)));
		mesh->material->texture = tex;
	}
	return mesh;
}

int main(int argc, char** argv)
{
	float d = 700, wx = 0, wz = 40, yaw = 0, tilt = 20;
	int n = 10, sizew = 1920, sizeh = 0, objects = 20, rings = 200, segments = 200;
	bool usetex = false, nolight = false, saving = false;
	for (int i = 1; i < argc; i++)
	{
		const std::string a = argv[i];
		const bool more = i + 1 < argc;
		if (a == "-d" && more) d = (float)atof(argv[++i]);
		else if (a == "-n" && more) n = atoi(argv[++i]);
		else if (a == "-w" && more) sizew = atoi(argv[++i]);
		else if (a == "-h" && more) sizeh = atoi(argv[++i]);
		else if (a == "-rx" && more) wx = (float)atof(argv[++i]);
		else if (a == "-rz" && more) wz = (float)atof(argv[++i]);
		else if (a == "-yaw" && more) yaw = (float)atof(argv[++i]);
		else if (a == "-tilt" && more) tilt = (float)atof(argv[++i]);
		else if (a == "-objects" && more) objects = atoi(argv[++i]);
		else if (a == "-m" && more) rings = atoi(argv[++i]);
		else if (a == "-seg" && more) segments = atoi(argv[++i]);
		else if (a == "-tex") usetex = true;
		else if (a == "-dark") nolight = true;
		else if (a == "-save") saving = true;
	}
	if (sizeh <= 0)
		sizeh = sizew * 9 / 16;
	wx = deg2rad(wx); wz = deg2rad(wz); yaw = deg2rad(yaw); tilt = deg2rad(tilt);
	const float fov = deg2rad(35.f);

	Shared<Scene> scene = new Scene();
	for (int i = 0; i < objects; i++)
	{
		Shared<TriMesh> shape = createObject(rings, segments, usetex);
		shape->material->diffuse = Vec3(rnd(1.f), rnd(1.f), rnd(1.f));
		shape->transform = Matrix4::translate(rnd(-180.f, 180.f), rnd(-180.f, 180.f), rnd(-100.f, 100.f)) *
		                   Matrix4::rotate(Vec3(rnd(1.f), rnd(1.f), rnd(1.f)));
		scene->children << Shared<SceneNode>(shape);
	}
	scene->ambientLight = 0.2f;

	Renderer renderer;
	renderer.setLight(Vec3(-0.4f, .6f, 1.f));
	renderer.setScene(scene);
	renderer.setSize(sizew, sizeh);
	renderer.setProjection(projectionFrustum(fov, renderer.aspect(), 10, 7000));
	renderer.setLighting(!nolight);
	renderer.setTexturing(usetex);
	renderer.setSaveNormals(false);

	float rx = -(float)PI / 2 + tilt, rz = yaw;
	long long covered = 0;
	const double t2 = nowSeconds();
	for (int i = 0; i < n; i++)
	{
		const float dt = 0.1f;
		rz += wz * dt;
		rx += wx * dt;
		renderer.setView(Matrix4::translate(0, 0, -d) * Matrix4::rotateX(rx) * Matrix4::rotateZ(rz));
		renderer.render();
#ifndef BENCH_SCENE_NO_IO
		if (saving)
		{
			char name[64];
			snprintf(name, sizeof(name), "bench%04i.ppm", i);
			savePPM(renderer.getImage(), name);
		}
#endif
	}
	// render() is asynchronous in the GPU build: reading a buffer completes the last frame
	const Array2<float> depth = renderer.getDepth();
	const double t6 = nowSeconds();
	for (int i = 0; i < depth.rows(); i++)
		for (int j = 0; j < depth.cols(); j++)
			covered += depth(i, j) < 1e10f;
	printf("t = %.3f (t frame = %.3f)\n", t6 - t2, (t6 - t2) / n);
	printf("covered pixels in the last frame: %lld of %d\n", covered, sizew * sizeh);
	return 0;
}
