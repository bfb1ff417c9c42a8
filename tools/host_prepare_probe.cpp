// Host time of Renderer::prepare() for a 10 000-mesh scene (no GPU needed): the numbers quoted in DESIGN.md §1 for the
// host thread pool. Build and run from the repository root:
//   g++ -std=c++11 -O3 -ffp-contract=off -Iinclude -Ithird_party/asl_shim tools/host_prepare_probe.cpp -o /tmp/host_prepare_probe \
//       -Lminirender_b200/lib -lminirender_b200 -Wl,-rpath,$PWD/minirender_b200/lib
//   /tmp/host_prepare_probe ; MINIRENDER_B200_HOST_THREADS=1 /tmp/host_prepare_probe
#include <minirender/Scene.h>
#include <minirender/Renderer.h>
#include <minirender/primitives.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>

using namespace minirender;
using namespace asl;

static double seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
	const int groups = argc > 1 ? atoi(argv[1]) : 100, perGroup = argc > 2 ? atoi(argv[2]) : 100, frames = argc > 3 ? atoi(argv[3]) : 300;
	Shared<Scene> scene = new Scene;
	for (int g = 0; g < groups; g++)
	{
		Shared<SceneNode> group = new SceneNode;
		group->transform = Matrix4::translate(g * 3.f, 0, 0) * Matrix4::rotateX(0.1f * g);
		scene->children << group;
		for (int k = 0; k < perGroup; k++)
		{
			Shared<TriMesh> m = createSphere(5.f, 8, 12);
			m->transform = Matrix4::translate(0, k * 2.f, 0) * Matrix4::rotateY(0.3f * k) * Matrix4::scale(Vec3(1.f, 1.2f, 0.8f));
			m->material = new Material;
			group->children << Shared<SceneNode>(m);
		}
	}
	Renderer r;
	r.setScene(scene);
	r.setSize(640, 360);
	r.setView(Matrix4::translate(0, 0, -30) * Matrix4::rotateX(-1.2f));
	r.setProjection(projectionFrustum(0.6f, 640.f / 360, 10, 7000));
	for (int i = 0; i < 20; i++)
		r.prepare();
	for (int block = 0; block < 6; block++) // (the first blocks include the scheduler spreading the new threads over the cores)
	{
		const double t0 = seconds();
		for (int i = 0; i < frames; i++)
			r.prepare();
		printf("%d renderables: prepare() %.1f us per frame\n", groups * perGroup, (seconds() - t0) / frames * 1e6);
	}
	return 0;
}
