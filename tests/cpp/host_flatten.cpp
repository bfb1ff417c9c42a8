// Host-side check of Renderer::prepare() on a many-node scene (no GPU needed): the renderables it hands to the C ABI must
// be the ones SceneNode::collectShapes gives, in that order, with modelview = view * world and
// normal matrix = modelview.inverse().t() bit for bit (reference src/Scene.cpp:13-36, src/Renderer.cpp:337-338) -
// whether the library flattens on one thread or deals sub-trees to its host pool - and a node type of the application's
// must have its own collectShapes called, on the calling thread.
// usage: host_flatten [custom]     prints "OK <renderables> <frames> <threads of the process>" or a line starting with FAIL
#include <minirender/Scene.h>
#include <minirender/Renderer.h>
#include <minirender/primitives.h>
#include <minirender_b200.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace minirender;
using namespace asl;

static std::thread::id g_caller;
static int g_foreignCalls = 0, g_foreignOffThread = 0;

// An application's node: emits its children last to first (so a stock flatten of it would give another order).
struct Reversed : public SceneNode
{
	virtual void collectShapes(Array<Renderable>& list, const Matrix4& xform)
	{
		g_foreignCalls++;
		if (std::this_thread::get_id() != g_caller)
			g_foreignOffThread++;
		const Matrix4 world = xform * transform;
		for (int i = children.length() - 1; i >= 0; i--)
			children[i]->collectShapes(list, world);
	}
};

static int processThreads() // Linux: /proc/self/status
{
	int n = 0;
	char line[256];
	FILE* f = fopen("/proc/self/status", "r");
	while (f && fgets(line, sizeof(line), f))
		if (!strncmp(line, "Threads:", 8))
			n = atoi(line + 8);
	if (f)
		fclose(f);
	return n;
}

static unsigned g_seed = 12345u;
static float rnd() { g_seed = g_seed * 1664525u + 1013904223u; return (float)(g_seed >> 8) / 16777216.0f; }
static Matrix4 someTransform()
{
	return Matrix4::translate(rnd() * 80 - 40, rnd() * 80 - 40, rnd() * 80 - 40) * Matrix4::rotateX(rnd() * 3) * Matrix4::rotateY(rnd() * 3) *
	       Matrix4::scale(Vec3(0.5f + rnd(), 0.5f + rnd(), 0.5f + rnd()));
}

static int check(Renderer& r, Scene* scene, const Matrix4& view, int frame)
{
	r.prepare();
	const mr_frame* f = r.frameDesc();
	Array<Renderable> want;
	scene->collectShapes(want, Matrix4::identity());
	if (f->n_renderables != want.length())
	{
		printf("FAIL frame %d: %d renderables, collectShapes gives %d\n", frame, f->n_renderables, want.length());
		return 1;
	}
	const mr_scene_desc* sd = r.sceneDesc();
	for (int i = 0; i < want.length(); i++)
	{
		const Matrix4 mv = view * want[i].transform;
		const Matrix4 nm = mv.inverse().t();
		float a[12], b[12];
		for (int k = 0; k < 3; k++)
			for (int j = 0; j < 4; j++)
			{
				a[4 * k + j] = mv(k, j);
				b[4 * k + j] = nm(k, j);
			}
		const mr_renderable& e = f->renderables[i];
		if (memcmp(a, e.modelview, sizeof(a)) || memcmp(b, e.normalmat, sizeof(b)))
		{
			printf("FAIL frame %d: matrices of entry %d differ\n", frame, i);
			return 1;
		}
		if (e.mesh < 0 || e.mesh >= sd->n_meshes || sd->meshes[e.mesh].positions != (const float*)want[i].mesh->vertices.ptr())
		{
			printf("FAIL frame %d: entry %d is not the mesh collectShapes put there\n", frame, i);
			return 1;
		}
	}
	return 0;
}

int main(int argc, char** argv)
{
	const bool custom = argc > 1 && !strcmp(argv[1], "custom");
	g_caller = std::this_thread::get_id();
	Shared<Scene> scene = new Scene;
	scene->transform = Matrix4::rotateZ(0.2f);
	Array<Shared<TriMesh> > shared;
	for (int i = 0; i < 5; i++)
		shared << createSphere(3.0f + i, 5, 6);
	Shared<SceneNode> lastGroup;
	int meshes = 0;
	for (int g = 0; g < 37; g++)
	{
		Shared<SceneNode> group = (custom && g % 9 == 4) ? Shared<SceneNode>(new Reversed) : Shared<SceneNode>(new SceneNode);
		group->transform = someTransform();
		scene->children << group;
		lastGroup = group;
		for (int k = 0; k < 40; k++)
		{
			Shared<TriMesh> m = createCube(1.0f + rnd());
			m->transform = someTransform();
			m->material = new Material;
			group->children << Shared<SceneNode>(m);
			meshes++;
			if (k % 7 == 3) // a mesh with children of its own: it comes before them
				for (int c = 0; c < 3; c++)
				{
					Shared<TriMesh> ch = createCube(0.5f);
					ch->transform = someTransform();
					m->children << Shared<SceneNode>(ch);
					meshes++;
				}
			if (k % 11 == 5) // the same mesh object in several places
			{
				group->children << Shared<SceneNode>(shared[(g + k) % 5]);
				meshes++;
			}
			if (k % 13 == 6) // a nested group
			{
				Shared<SceneNode> inner = new SceneNode;
				inner->transform = someTransform();
				for (int c = 0; c < 4; c++)
				{
					Shared<TriMesh> ch = createSphere(1.0f, 4, 5);
					ch->transform = someTransform();
					inner->children << Shared<SceneNode>(ch);
					meshes++;
				}
				group->children << inner;
			}
		}
	}
	// a mesh directly under the scene, between the groups and behind them
	scene->children << Shared<SceneNode>(shared[0]);
	meshes++;

	Renderer r;
	r.setScene(scene);
	r.setSize(320, 200);
	Matrix4 view = Matrix4::translate(0, 0, -150) * Matrix4::rotateX(-0.4f);
	r.setView(view);
	r.setProjection(projectionFrustum(0.7f, 1.6f, 1.0f, 1000.0f));
	int bad = 0, frames = 0;
	for (int frame = 0; frame < 6 && !bad; frame++, frames++)
	{
		if (frame == 2) // transforms move
		{
			scene->children[3]->transform = someTransform();
			view = Matrix4::translate(1, 2, -140) * Matrix4::rotateY(0.3f);
			r.setView(view);
		}
		if (frame == 3) // the structure changes: a node more, a node less
		{
			Shared<TriMesh> extra = createCube(2.0f);
			extra->transform = someTransform();
			scene->children[5]->children << Shared<SceneNode>(extra);
			meshes++;
		}
		if (frame == 4)
		{
			Shared<SceneNode> g = new SceneNode; // an empty group, and a group holding only an empty group
			g->children << Shared<SceneNode>(new SceneNode);
			scene->children << g;
		}
		bad = check(r, scene, view, frame);
	}
	if (!bad && custom && (g_foreignCalls == 0 || g_foreignOffThread != 0))
	{
		printf("FAIL application node: %d calls, %d of them off the calling thread\n", g_foreignCalls, g_foreignOffThread);
		bad = 1;
	}
	if (!bad)
		printf("OK %d %d %d\n", r.frameDesc()->n_renderables, frames, processThreads());
	return bad;
}
