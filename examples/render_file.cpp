// render_file: load a model (STL / OBJ / X3D), fit the camera, render one or many turntable frames
// and save them as PPM. The batch part of the reference's `render` tool (reference samples/main.cpp:
// same camera fit :118-127, same light, projection and view sequence :129-158) without its console
// painter, written against the public minirender C++ API only, so it builds unchanged against this
// repo's headers + libminirender_b200.so (frames rendered on the GPU).
//
//   render_file model.obj [-n frames] [-w width] [-h height] [-d distance] [-rz deg/s] [-rx deg/s]
//               [-yaw deg] [-tilt deg] [-fov deg] [-yup] [-o out%04i.ppm] [-silent]
#include <minirender/Renderer.h>
#include <minirender/Scene.h>
#include <minirender/io.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace asl;
using namespace minirender;

static double nowSeconds()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv)
{
	const char* model = 0;
	int n = 1, sizew = 800, sizeh = 0;
	float d = 140, wx = 0, wz = 40, yaw = 0, tilt = 20, fovDeg = 35;
	bool fit = true, yup = false, silent = false;
	std::string outname;
	for (int i = 1; i < argc; i++)
	{
		const std::string a = argv[i];
		const bool more = i + 1 < argc;
		if (a == "-n" && more) n = atoi(argv[++i]);
		else if (a == "-w" && more) sizew = atoi(argv[++i]);
		else if (a == "-h" && more) sizeh = atoi(argv[++i]);
		else if (a == "-d" && more) { d = (float)atof(argv[++i]); fit = false; }
		else if (a == "-rx" && more) wx = (float)atof(argv[++i]);
		else if (a == "-rz" && more) wz = (float)atof(argv[++i]);
		else if (a == "-yaw" && more) yaw = (float)atof(argv[++i]);
		else if (a == "-tilt" && more) tilt = (float)atof(argv[++i]);
		else if (a == "-fov" && more) fovDeg = (float)atof(argv[++i]);
		else if (a == "-o" && more) outname = argv[++i];
		else if (a == "-yup") yup = true;
		else if (a == "-fit") fit = true;
		else if (a == "-silent") silent = true;
		else if (a[0] != '-') model = argv[i];
	}
	if (!model)
	{
		printf("usage: render_file model.(stl|obj|x3d) [-n frames] [-w W] [-h H] [-d dist] [-rz deg/s] [-rx deg/s] [-yaw deg] [-tilt deg] [-fov deg] [-yup] [-o out%%04i.ppm]\n");
		return 0;
	}
	if (sizeh <= 0)
		sizeh = sizew * 3 / 4;
	if (outname.empty())
		outname = n == 1 ? "out.ppm" : "out%04i.ppm";
	const float fov = deg2rad(fovDeg);
	wx = deg2rad(wx); wz = deg2rad(wz); yaw = deg2rad(yaw); tilt = deg2rad(tilt);

	const double t1 = nowSeconds();
	Shared<SceneNode> shape = loadMesh(model);
	if (!shape)
	{
		printf("Cannot load model\n");
		return 1;
	}
	const double t2 = nowSeconds();
	if (!silent)
		printf("load %s %.3f s\n", model, t2 - t1);

	Shared<Scene> scene = new Scene();
	if (yup)
		scene->transform = Matrix4::rotateX(PIf / 2);
	scene->children << shape;
	scene->ambientLight = 0.2f;
	const BBox box = scene->getBbox();
	const Vec3 size = box.size(), center = box.center();
	if (fit)
	{
		const float dh = max(size.x, size.y) / (2 * (float)tan(fov * sizew / sizeh / 2));
		const float dv = size.z / (2 * (float)tan(fov / 2));
		d = 1.55f * max(dh, dv);
		scene->transform = Matrix4::translate(-center) * scene->transform;
	}

	Renderer renderer;
	renderer.setBackground(Vec3(0, 0, 0));
	renderer.setLight(Vec3(-0.3f, 0.55f, 1));
	renderer.setScene(scene);
	renderer.setSize(sizew, sizeh);
	renderer.setProjection(projectionFrustum(fov, renderer.aspect(), 10, 7000));
	renderer.setTexturing(true);

	float rx = -(float)PI / 2 + tilt, rz = yaw;
	double tsave = 0;
	for (int i = 0; i < n; i++)
	{
		const float dt = 0.1f;
		rz += wz * dt;
		rx += wx * dt;
		renderer.setView(Matrix4::translate(0, 0, -d) * Matrix4::rotateX(rx) * Matrix4::rotateZ(rz));
		renderer.render();
		const double ta = nowSeconds();
		char name[1024];
		snprintf(name, sizeof(name), outname.c_str(), i);
		savePPM(renderer.getImage(), n == 1 ? outname.c_str() : name);
		tsave += nowSeconds() - ta;
	}
	const double t6 = nowSeconds();
	if (!silent)
	{
		printf("t = %.3f (t frame = %.3f)\n", t6 - t2, (t6 - t2) / n);
		printf("t read+save = %.3f\n", tsave / n);
	}
	return 0;
}
