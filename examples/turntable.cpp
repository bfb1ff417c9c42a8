// Drop-in check: this file uses only the public minirender C++ API (Scene / TriMesh / Material /
// Renderer / primitives / projection builders), the way samples/bench.cpp of the reference does.
// It compiles unchanged against
//   - this repo's headers + libminirender_b200.so  (render() runs on the GPU), and
//   - the reference's own headers + sources        (CPU; built by tests for comparison),
// renders a small turntable, and prints an FNV-1a hash of the depth buffer bits and of the 8-bit
// quantised image (savePPM's rule) per frame. Equal depth hashes = bit-exact coverage and depth.
#include <minirender/Renderer.h>
#include <minirender/Scene.h>
#include <minirender/primitives.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>

using namespace asl;
using namespace minirender;

static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull)
{
	const unsigned char* b = (const unsigned char*)p;
	for (size_t i = 0; i < n; i++)
		h = (h ^ b[i]) * 1099511628211ull;
	return h;
}

int main(int argc, char** argv)
{
	const int frames = argc > 1 ? atoi(argv[1]) : 3;
	const int w = argc > 2 ? atoi(argv[2]) : 320, h = argc > 3 ? atoi(argv[3]) : 200;

	Shared<Scene> scene = new Scene();
	scene->ambientLight = 0.2f;
	Shared<SceneNode> group = new SceneNode();
	group->transform = Matrix4::translate(0, 0, 10) * Matrix4::rotate(Vec3(0.2f, 0.1f, 0.3f));
	Shared<TriMesh> ball = createSphere(40.0f, 24, 48);
	ball->material->diffuse = Vec3(0.9f, 0.4f, 0.2f);
	ball->material->shininess = 20;
	ball->transform = Matrix4::translate(-30, 0, 0);
	Shared<TriMesh> box = createCube(45.0f);
	box->transform = Matrix4::translate(45, 10, -5) * Matrix4::rotate(Vec3(0.5f, 0.7f, 0.1f));
	Shared<TriMesh> pipe = createCylinder(12.0f, 90.0f, 20, 2);
	pipe->material->emissive = Vec3(0.05f, 0.0f, 0.1f);
	pipe->transform = Matrix4::translate(0, -40, 20) * Matrix4::rotateX(1.1f);
	group->children << ball << box;
	box->children << pipe; // a mesh with a child: flattened after its parent
	scene->add(group);
	scene->add(ball); // the same mesh object a second time, under the root

	Renderer renderer;
	renderer.setScene(scene);
	renderer.setSize(w, h);
	renderer.setLight(Vec3(-0.4f, 0.6f, 1.0f));
	renderer.setSaveNormals(false);
	renderer.setBackground(Vec3(0.1f, 0.1f, 0.15f));
	renderer.setProjection(projectionFrustum(deg2rad(35.0f), renderer.aspect(), 10, 3000));

	for (int i = 0; i < frames; i++)
	{
		renderer.setView(Matrix4::translate(0, 0, -260) * Matrix4::rotateX(-1.0f) * Matrix4::rotateZ(0.35f * i));
		ball->transform = Matrix4::translate(-30 + 5.0f * i, 0, 0); // users move nodes between frames
		renderer.render();
		Array2<Vec3> image = renderer.getImage();
		Array2<float> depth = renderer.getDepth();
		int covered = 0;
		uint64_t hi = 1469598103934665603ull;
		for (int r = 0; r < image.rows(); r++)
			for (int c = 0; c < image.cols(); c++)
			{
				covered += depth(r, c) < 1e10f;
				Vec3 v = image(r, c) * 255.0f;
				unsigned char q[3] = { (unsigned char)clamp(v.x, 0.0f, 255.0f), (unsigned char)clamp(v.y, 0.0f, 255.0f),
					                   (unsigned char)clamp(v.z, 0.0f, 255.0f) };
				hi = fnv(q, 3, hi);
			}
		printf("frame %d covered %d depth %016llx rgb8 %016llx\n", i, covered,
		       (unsigned long long)fnv(&depth(0, 0), sizeof(float) * (size_t)w * h), (unsigned long long)hi);
	}
	return 0;
}
