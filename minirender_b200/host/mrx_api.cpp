// Flat C view of the minirender C++ API; see mrx_api.h. Compiles unchanged against this repo's
// headers (define MRX_PRODUCT) or against the reference's headers (oracle/_ref build).
#include "mrx_api.h"

#include <minirender/Renderer.h>
#include <minirender/Scene.h>
#include <minirender/primitives.h>
#include <minirender/io.h>
#ifdef MRX_PRODUCT
#include <minirender_b200.h>
#endif

#include <cstring>
#include <exception>
#include <string>
#include <vector>

using namespace asl;
using namespace minirender;

namespace {

thread_local std::string g_error;

struct SceneBox
{
	Shared<Scene> scene;
	std::vector<Shared<SceneNode> > nodes;
	std::vector<Shared<Material> > materials;
};

Matrix4 toMatrix(const float* m)
{
	if (!m)
		return Matrix4::identity();
	return Matrix4(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
}

void fromMatrix(float* out, const Matrix4& m)
{
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++)
			out[4 * i + j] = m(i, j);
}

int attach(SceneBox* sb, int parent, const Shared<SceneNode>& node)
{
	if (parent < 0)
		sb->scene->children << node;
	else if (parent < (int)sb->nodes.size())
		sb->nodes[parent]->children << node;
	else
	{
		g_error = "bad parent node id";
		return -1;
	}
	return 0;
}

TriMesh* meshOf(SceneBox* sb, int node)
{
	if (node < 0 || node >= (int)sb->nodes.size())
		return 0;
	return dynamic_cast<TriMesh*>((SceneNode*)sb->nodes[node]);
}

int64_t countTriangles(SceneNode* n)
{
	int64_t t = 0;
	if (TriMesh* m = dynamic_cast<TriMesh*>(n))
		t += m->indices.length() / 3;
	for (int i = 0; i < n->children.length(); i++)
		t += countTriangles(n->children[i]);
	return t;
}

}

#define MRX_TRY try {
#define MRX_CATCH(ret)                                                                                             \
	}                                                                                                              \
	catch (const std::exception& e) { g_error = e.what(); return ret; }                                            \
	catch (...) { g_error = "unknown C++ exception"; return ret; }

extern "C" {

const char* mrx_last_error(void) { return g_error.c_str(); }

const char* mrx_backend(void)
{
#ifdef MRX_PRODUCT
	return "b200";
#else
	return "reference";
#endif
}

void* mrx_scene_new(void)
{
	MRX_TRY
	SceneBox* sb = new SceneBox;
	sb->scene = new Scene();
	return sb;
	MRX_CATCH(0)
}

void mrx_scene_free(void* scene) { delete (SceneBox*)scene; }

void mrx_scene_set_ambient(void* scene, float ambient) { ((SceneBox*)scene)->scene->ambientLight = ambient; }

int mrx_add_material(void* scene, const float* diffuse, const float* specular, const float* emissive, float shininess,
                     const float* texels, int rows, int cols)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	Shared<Material> m = new Material();
	if (diffuse) m->diffuse = Vec3(diffuse[0], diffuse[1], diffuse[2]);
	if (specular) m->specular = Vec3(specular[0], specular[1], specular[2]);
	if (emissive) m->emissive = Vec3(emissive[0], emissive[1], emissive[2]);
	m->shininess = shininess;
	if (texels && rows > 0 && cols > 0)
	{
		Array2<Vec3> tex(rows, cols);
		for (int i = 0; i < rows; i++)
			for (int j = 0; j < cols; j++)
			{
				const float* t = texels + 3 * ((size_t)i * cols + j);
				tex(i, j) = Vec3(t[0], t[1], t[2]);
			}
		m->texture = tex;
	}
	sb->materials.push_back(m);
	return (int)sb->materials.size() - 1;
	MRX_CATCH(-1)
}

int mrx_material_update(void* scene, int material, const float* diffuse, const float* specular, const float* emissive, float shininess)
{
	SceneBox* sb = (SceneBox*)scene;
	if (material < 0 || material >= (int)sb->materials.size())
		return -1;
	Material* m = sb->materials[material];
	if (diffuse) m->diffuse = Vec3(diffuse[0], diffuse[1], diffuse[2]);
	if (specular) m->specular = Vec3(specular[0], specular[1], specular[2]);
	if (emissive) m->emissive = Vec3(emissive[0], emissive[1], emissive[2]);
	m->shininess = shininess;
	return 0;
}

int mrx_add_group(void* scene, int parent, const float* xf)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	Shared<SceneNode> node = new SceneNode();
	node->transform = toMatrix(xf);
	if (attach(sb, parent, node))
		return -1;
	sb->nodes.push_back(node);
	return (int)sb->nodes.size() - 1;
	MRX_CATCH(-1)
}

int mrx_add_mesh(void* scene, int parent, const float* xf, const float* pos, int npos, const float* nrm, int nnrm, const float* uv,
                 int nuv, const int32_t* ipos, const int32_t* inrm, const int32_t* iuv, int ntri, int material)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	Shared<TriMesh> mesh = new TriMesh();
	mesh->transform = toMatrix(xf);
	mesh->vertices.resize(npos);
	for (int i = 0; i < npos; i++)
		mesh->vertices[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
	mesh->normals.resize(nnrm);
	for (int i = 0; i < nnrm; i++)
		mesh->normals[i] = Vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
	if (uv && nuv > 0)
	{
		mesh->texcoords.resize(nuv);
		for (int i = 0; i < nuv; i++)
			mesh->texcoords[i] = Vec2(uv[2 * i], uv[2 * i + 1]);
	}
	mesh->indices.resize(3 * ntri);
	mesh->normalsI.resize(3 * ntri);
	for (int i = 0; i < 3 * ntri; i++)
	{
		mesh->indices[i] = ipos[i];
		mesh->normalsI[i] = inrm ? inrm[i] : ipos[i];
	}
	if (iuv)
	{
		mesh->texcoordsI.resize(3 * ntri);
		for (int i = 0; i < 3 * ntri; i++)
			mesh->texcoordsI[i] = iuv[i];
	}
	if (material >= 0)
	{
		if (material >= (int)sb->materials.size())
		{
			g_error = "bad material id";
			return -1;
		}
		mesh->material = sb->materials[material];
	}
	if (attach(sb, parent, mesh))
		return -1;
	sb->nodes.push_back(mesh);
	return (int)sb->nodes.size() - 1;
	MRX_CATCH(-1)
}

int mrx_add_primitive(void* scene, int parent, const float* xf, int kind, float a, float b, int n1, int n2, int caps, int material,
                      int with_uv_index)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	Shared<TriMesh> mesh;
	switch (kind)
	{
	case 0: mesh = createCube(a); break;
	case 1: mesh = createCylinder(a, b, n1, n2, caps != 0); break;
	case 2: mesh = createSphere(a, n1, n2); break;
	default: g_error = "bad primitive kind"; return -1;
	}
	mesh->transform = toMatrix(xf);
	if (with_uv_index)
		mesh->texcoordsI = mesh->indices;
	if (material >= 0)
	{
		if (material >= (int)sb->materials.size())
		{
			g_error = "bad material id";
			return -1;
		}
		mesh->material = sb->materials[material];
	}
	if (attach(sb, parent, mesh))
		return -1;
	sb->nodes.push_back(mesh);
	return (int)sb->nodes.size() - 1;
	MRX_CATCH(-1)
}

int mrx_add_instance(void* scene, int parent, int node)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	if (node < 0 || node >= (int)sb->nodes.size())
	{
		g_error = "bad node id";
		return -1;
	}
	return attach(sb, parent, sb->nodes[node]);
	MRX_CATCH(-1)
}

int mrx_node_set_transform(void* scene, int node, const float* xf)
{
	SceneBox* sb = (SceneBox*)scene;
	if (node < 0 || node >= (int)sb->nodes.size())
		return -1;
	sb->nodes[node]->transform = toMatrix(xf);
	return 0;
}

int mrx_mesh_counts(void* scene, int node, int32_t* counts)
{
	TriMesh* m = meshOf((SceneBox*)scene, node);
	if (!m)
		return -1;
	counts[0] = m->vertices.length();
	counts[1] = m->normals.length();
	counts[2] = m->texcoords.length();
	counts[3] = m->indices.length();
	counts[4] = m->normalsI.length();
	counts[5] = m->texcoordsI.length();
	return 0;
}

int mrx_mesh_copy(void* scene, int node, float* pos, float* nrm, float* uv, int32_t* ipos, int32_t* inrm, int32_t* iuv)
{
	TriMesh* m = meshOf((SceneBox*)scene, node);
	if (!m)
		return -1;
	if (pos) memcpy(pos, m->vertices.ptr(), sizeof(Vec3) * m->vertices.length());
	if (nrm) memcpy(nrm, m->normals.ptr(), sizeof(Vec3) * m->normals.length());
	if (uv) memcpy(uv, m->texcoords.ptr(), sizeof(Vec2) * m->texcoords.length());
	if (ipos) memcpy(ipos, m->indices.ptr(), sizeof(int) * m->indices.length());
	if (inrm) memcpy(inrm, m->normalsI.ptr(), sizeof(int) * m->normalsI.length());
	if (iuv) memcpy(iuv, m->texcoordsI.ptr(), sizeof(int) * m->texcoordsI.length());
	return 0;
}

int mrx_scene_bbox(void* scene, float* out6)
{
	MRX_TRY
	BBox b = ((SceneBox*)scene)->scene->getBbox();
	out6[0] = b.pmin.x; out6[1] = b.pmin.y; out6[2] = b.pmin.z;
	out6[3] = b.pmax.x; out6[4] = b.pmax.y; out6[5] = b.pmax.z;
	return 0;
	MRX_CATCH(-1)
}

int64_t mrx_scene_triangles(void* scene)
{
	return countTriangles((SceneNode*)((SceneBox*)scene)->scene);
}

// ---- renderer ----

void* mrx_renderer_new(void)
{
	MRX_TRY
	Renderer* r = new Renderer();
	r->setSaveNormals(false); // the reference leaves this member uninitialised
	r->setView(Matrix4::identity());
	return r;
	MRX_CATCH(0)
}

void mrx_renderer_free(void* r) { delete (Renderer*)r; }

int mrx_renderer_set_scene(void* r, void* scene)
{
	MRX_TRY
	((Renderer*)r)->setScene(((SceneBox*)scene)->scene);
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_set_size(void* r, int w, int h)
{
	MRX_TRY
	((Renderer*)r)->setSize(w, h);
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_set_projection(void* r, const float* m16) { ((Renderer*)r)->setProjection(toMatrix(m16)); return 0; }
int mrx_renderer_set_view(void* r, const float* m16) { ((Renderer*)r)->setView(toMatrix(m16)); return 0; }
int mrx_renderer_set_light(void* r, const float* v, int point) { ((Renderer*)r)->setLight(Vec3(v[0], v[1], v[2]), point != 0); return 0; }
int mrx_renderer_set_lighting(void* r, int on) { ((Renderer*)r)->setLighting(on != 0); return 0; }
int mrx_renderer_set_texturing(void* r, int on) { ((Renderer*)r)->setTexturing(on != 0); return 0; }
int mrx_renderer_set_save_normals(void* r, int on) { ((Renderer*)r)->setSaveNormals(on != 0); return 0; }
int mrx_renderer_set_background(void* r, const float* c) { ((Renderer*)r)->setBackground(Vec3(c[0], c[1], c[2])); return 0; }

int mrx_renderer_clear(void* r)
{
	MRX_TRY
	((Renderer*)r)->clear();
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_render(void* r)
{
	MRX_TRY
	((Renderer*)r)->render();
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_paint_mesh(void* r, void* scene, int node, const float* xf)
{
	MRX_TRY
	TriMesh* m = meshOf((SceneBox*)scene, node);
	if (!m)
	{
		g_error = "node is not a mesh";
		return -1;
	}
	((Renderer*)r)->paintMesh(m, toMatrix(xf));
	return 0;
	MRX_CATCH(-1)
}

int mrx_mesh_apply_transform(void* scene, int node)
{
	MRX_TRY
	TriMesh* m = meshOf((SceneBox*)scene, node);
	if (!m)
	{
		g_error = "node is not a mesh";
		return -1;
	}
	m->applyTransform();
	return 0;
	MRX_CATCH(-1)
}

int mrx_mesh_move_vertex(void* scene, int node, int i, float dx, float dy, float dz)
{
	MRX_TRY
	TriMesh* m = meshOf((SceneBox*)scene, node);
	if (!m || i < 0 || i >= m->vertices.length())
	{
		g_error = "node is not a mesh or vertex index out of range";
		return -1;
	}
	m->vertices[i] = m->vertices[i] + Vec3(dx, dy, dz);
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_paint_triangle(void* r, const float* v, int world)
{
	MRX_TRY
	Vertex c[3];
	for (int i = 0; i < 3; i++)
	{
		const float* p = v + 8 * i;
		c[i].position = Vec3(p[0], p[1], p[2]);
		c[i].normal = Vec3(p[3], p[4], p[5]);
		c[i].uv = Vec2(p[6], p[7]);
	}
	((Renderer*)r)->paintTriangle(c[0], c[1], c[2], world != 0);
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_set_material(void* r, void* scene, int material)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	if (material < 0 || material >= (int)sb->materials.size())
	{
		g_error = "material index out of range";
		return -1;
	}
	((Renderer*)r)->setMaterial(sb->materials[material]);
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_get_image(void* r, float* out)
{
	MRX_TRY
	Array2<Vec3> img = ((Renderer*)r)->getImage();
	memcpy(out, &img(0, 0), sizeof(Vec3) * (size_t)img.rows() * img.cols());
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_get_depth(void* r, float* out)
{
	MRX_TRY
	Array2<float> d = ((Renderer*)r)->getDepth();
	memcpy(out, &d(0, 0), sizeof(float) * (size_t)d.rows() * d.cols());
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_get_normals(void* r, float* out)
{
	MRX_TRY
	Array2<Vec3> n = ((Renderer*)r)->getNormalsImage();
	memcpy(out, &n(0, 0), sizeof(Vec3) * (size_t)n.rows() * n.cols());
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_get_range(void* r, float* out)
{
	MRX_TRY
	Array2<Vec3> p = ((Renderer*)r)->getRangeImage();
	memcpy(out, &p(0, 0), sizeof(Vec3) * (size_t)p.rows() * p.cols());
	return 0;
	MRX_CATCH(-1)
}

int mrx_quantize_rgb8(const float* image, int w, int h, uint8_t* out)
{
	// savePPM's rule (reference src/io.cpp:358-361) through the asl types it is written in
	const size_t n = (size_t)w * h;
	for (size_t i = 0; i < n; i++)
	{
		Vec3 value = Vec3(image[3 * i], image[3 * i + 1], image[3 * i + 2]) * 255.0f;
		out[3 * i] = (byte)clamp(value.x, 0.0f, 255.0f);
		out[3 * i + 1] = (byte)clamp(value.y, 0.0f, 255.0f);
		out[3 * i + 2] = (byte)clamp(value.z, 0.0f, 255.0f);
	}
	return 0;
}

// ---- matrices ----

void mrx_mat_translate(float* out, float x, float y, float z) { fromMatrix(out, Matrix4::translate(x, y, z)); }
void mrx_mat_scale(float* out, float x, float y, float z) { fromMatrix(out, Matrix4::scale(Vec3(x, y, z))); }
void mrx_mat_rotate_x(float* out, float a) { fromMatrix(out, Matrix4::rotateX(a)); }
void mrx_mat_rotate_y(float* out, float a) { fromMatrix(out, Matrix4::rotateY(a)); }
void mrx_mat_rotate_z(float* out, float a) { fromMatrix(out, Matrix4::rotateZ(a)); }
void mrx_mat_rotate_axis(float* out, float x, float y, float z, float angle) { fromMatrix(out, Matrix4::rotate(Vec3(x, y, z), angle)); }
void mrx_mat_rotate_vec(float* out, float x, float y, float z) { fromMatrix(out, Matrix4::rotate(Vec3(x, y, z))); }
void mrx_mat_mul(float* out, const float* a, const float* b) { fromMatrix(out, toMatrix(a) * toMatrix(b)); }
void mrx_mat_inverse(float* out, const float* a) { fromMatrix(out, toMatrix(a).inverse()); }

int mrx_projection(float* out, int kind, const float* p)
{
	switch (kind)
	{
	case 0: fromMatrix(out, projectionOrtho(p[0], p[1], p[2], p[3], p[4], p[5])); return 0;
	case 1: fromMatrix(out, projectionPerspective(p[0], p[1], p[2], p[3], p[4], p[5])); return 0;
	case 2: fromMatrix(out, projectionFrustum(p[0], p[1], p[2], p[3])); return 0;
	case 3: fromMatrix(out, projectionFrustumH(p[0], p[1], p[2], p[3])); return 0;
	case 4: fromMatrix(out, projectionOrtho(p[0], p[1], p[2], p[3])); return 0;
	}
	g_error = "bad projection kind";
	return -1;
}

void mrx_projection_cv(float* out, const float* K16, float w, float h, float n, float f)
{
	fromMatrix(out, projectionCV(toMatrix(K16), w, h, n, f));
}

// ---- product-only entry points ----
#ifdef MRX_PRODUCT

int mrx_renderer_set_device(void* r, int device)
{
	MRX_TRY
	((Renderer*)r)->setDevice(device);
	return 0;
	MRX_CATCH(-1)
}

int mrx_renderer_set_row_range(void* r, int begin, int end) { ((Renderer*)r)->setRowRange(begin, end); return 0; }
int mrx_renderer_invalidate_geometry(void* r) { ((Renderer*)r)->invalidateGeometry(); return 0; }

int mrx_renderer_prepare(void* r)
{
	MRX_TRY
	((Renderer*)r)->prepare();
	return 0;
	MRX_CATCH(-1)
}

const void* mrx_renderer_scene_desc(void* r) { return ((Renderer*)r)->sceneDesc(); }
const void* mrx_renderer_frame_desc(void* r) { return ((Renderer*)r)->frameDesc(); }

void* mrx_renderer_context(void* r)
{
	MRX_TRY
	return ((Renderer*)r)->context();
	MRX_CATCH(0)
}

int mrx_renderer_get_rgb8(void* r, uint8_t* out)
{
	MRX_TRY
	Array<byte> a = ((Renderer*)r)->getImageRGB8();
	memcpy(out, a.ptr(), (size_t)a.length());
	return 0;
	MRX_CATCH(-1)
}

const float* mrx_renderer_image_ptr(void* r)
{
	MRX_TRY
	Array2<Vec3> img = ((Renderer*)r)->getImage();
	return (const float*)&img(0, 0);
	MRX_CATCH(0)
}

const float* mrx_renderer_depth_ptr(void* r)
{
	MRX_TRY
	Array2<float> d = ((Renderer*)r)->getDepth();
	return &d(0, 0);
	MRX_CATCH(0)
}

int mrx_renderer_synchronize(void* r)
{
	MRX_TRY
	((Renderer*)r)->synchronize();
	return 0;
	MRX_CATCH(-1)
}

#endif // MRX_PRODUCT

// ---- file formats (the reference's include/minirender/io.h): the same calls on either build ----

int mrx_save_ppm(const float* image, int w, int h, const char* filename)
{
	MRX_TRY
	Array2<Vec3> img(h, w);
	memcpy(&img(0, 0), image, sizeof(Vec3) * (size_t)w * h);
	savePPM(img, filename);
	return 0;
	MRX_CATCH(-1)
}

int mrx_load_ppm(const char* filename, float* out, int* rows, int* cols)
{
	MRX_TRY
	Array2<Vec3> img = loadPPM(filename);
	*rows = img.rows();
	*cols = img.cols();
	if (img.rows() == 0)
		return -1;
	if (out)
		memcpy(out, &img(0, 0), sizeof(Vec3) * (size_t)img.rows() * img.cols());
	return 0;
	MRX_CATCH(-1)
}

// Registers a loaded subtree's nodes (pre-order) so that tests can address them by id.
static void registerTree(SceneBox* sb, const Shared<SceneNode>& n)
{
	sb->nodes.push_back(n);
	for (int i = 0; i < n->children.length(); i++)
		registerTree(sb, n->children[i]);
}

int mrx_scene_load(void* scene, int parent, const char* filename)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	Shared<SceneNode> node = loadMesh(filename);
	if (!node)
	{
		g_error = std::string("cannot load ") + filename;
		return -1;
	}
	if (attach(sb, parent, node))
		return -1;
	const int first = (int)sb->nodes.size();
	registerTree(sb, node);
	return first;
	MRX_CATCH(-1)
}

int mrx_scene_node_count(void* scene) { return (int)((SceneBox*)scene)->nodes.size(); }

int mrx_node_info(void* scene, int node, int32_t* is_mesh, int32_t* n_children, float* transform16)
{
	SceneBox* sb = (SceneBox*)scene;
	if (node < 0 || node >= (int)sb->nodes.size())
		return -1;
	SceneNode* n = sb->nodes[node];
	if (is_mesh) *is_mesh = dynamic_cast<TriMesh*>(n) ? 1 : 0;
	if (n_children) *n_children = n->children.length();
	if (transform16) fromMatrix(transform16, n->transform);
	return 0;
}

int mrx_mesh_material(void* scene, int node, float* out11, int32_t* tex_rows, int32_t* tex_cols)
{
	TriMesh* m = meshOf((SceneBox*)scene, node);
	if (!m || !m->material)
		return -1;
	const Material& a = *m->material;
	const float v[11] = { a.diffuse.x, a.diffuse.y, a.diffuse.z, a.specular.x, a.specular.y, a.specular.z,
		                  a.emissive.x, a.emissive.y, a.emissive.z, a.shininess, a.opacity };
	memcpy(out11, v, sizeof(v));
	if (tex_rows) *tex_rows = a.texture.rows();
	if (tex_cols) *tex_cols = a.texture.cols();
	return 0;
}

int mrx_save_stl(void* scene, int node, const char* filename)
{
	MRX_TRY
	SceneBox* sb = (SceneBox*)scene;
	if (!meshOf(sb, node))
		return -1;
	Shared<TriMesh> mesh = sb->nodes[node].as<TriMesh>();
	saveSTL(mesh, filename);
	return 0;
	MRX_CATCH(-1)
}

int mrx_save_xyz(const float* points, int w, int h, const float* m16, const char* filename)
{
	MRX_TRY
	Array2<Vec3> pts(h, w);
	memcpy(&pts(0, 0), points, sizeof(Vec3) * (size_t)w * h);
	saveXYZ(pts, filename, toMatrix(m16));
	return 0;
	MRX_CATCH(-1)
}

#ifdef MRX_PRODUCT
// x3d.cpp's fan triangulation is a file-local function in the reference; the product exposes its own for the tests
int mrx_triangulate(const int32_t* in, int n, int32_t* out)
{
	Array<int> a;
	for (int i = 0; i < n; i++)
		a << in[i];
	const Array<int> t = triangulateIndices(a);
	if (out)
		memcpy(out, t.ptr(), sizeof(int) * t.length());
	return t.length();
}

#endif


}
