/*
 * mrx_api.h — flat C view of the minirender C++ API (Scene / TriMesh / Material / Renderer,
 * primitives, projection builders, asl::Matrix4 helpers) for ctypes.
 *
 * The same mrx_api.cpp is compiled twice:
 *   - against this repo's include/minirender/ headers into libminirender_b200.so (the product:
 *     render() runs the CUDA pipeline), and
 *   - against the reference's own headers and unmodified sources into
 *     oracle/_ref/libminirender_ref.so (the CPU oracle),
 * which is the drop-in claim in executable form: one client source, either implementation.
 * Functions marked [product only] exist only in the first build.
 */
#ifndef MRX_API_H
#define MRX_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MRX_API __attribute__((visibility("default")))

MRX_API const char* mrx_last_error(void);
MRX_API const char* mrx_backend(void); /* "b200" or "reference" */

/* ---- scene ---- */
MRX_API void* mrx_scene_new(void);
MRX_API void mrx_scene_free(void* scene);
MRX_API void mrx_scene_set_ambient(void* scene, float ambient);
/* texels: rows*cols*3 floats or NULL. Returns the material id. */
MRX_API int mrx_add_material(void* scene, const float* diffuse, const float* specular, const float* emissive, float shininess,
                             const float* texels, int rows, int cols);
MRX_API int mrx_material_update(void* scene, int material, const float* diffuse, const float* specular, const float* emissive, float shininess);
/* parent: node id or -1 for the scene root. xf: row-major 4x4 or NULL (identity). Returns the node id. */
MRX_API int mrx_add_group(void* scene, int parent, const float* xf);
MRX_API int mrx_add_mesh(void* scene, int parent, const float* xf, const float* pos, int npos, const float* nrm, int nnrm,
                         const float* uv, int nuv, const int32_t* ipos, const int32_t* inrm, const int32_t* iuv, int ntri,
                         int material /* -1: none (renderer default) */);
/* kind 0: cube(a=size)  1: cylinder(a=radius,b=height,n1=segments,n2=heightSegments,caps)  2: sphere(a=radius,n1=lat,n2=lon).
 * material -1 keeps the Material the generator attaches. with_uv_index != 0 sets texcoordsI = indices. */
MRX_API int mrx_add_primitive(void* scene, int parent, const float* xf, int kind, float a, float b, int n1, int n2, int caps,
                              int material, int with_uv_index);
MRX_API int mrx_add_instance(void* scene, int parent, int node); /* same node object under another parent */
MRX_API int mrx_node_set_transform(void* scene, int node, const float* xf);
/* counts[6] = n_positions, n_normals, n_texcoords, n_indices, n_normalsI, n_texcoordsI */
MRX_API int mrx_mesh_counts(void* scene, int node, int32_t* counts);
MRX_API int mrx_mesh_copy(void* scene, int node, float* pos, float* nrm, float* uv, int32_t* ipos, int32_t* inrm, int32_t* iuv);
MRX_API int mrx_scene_bbox(void* scene, float* out6);
MRX_API int64_t mrx_scene_triangles(void* scene); /* triangles submitted per frame (with instancing) */

/* ---- renderer (method-for-method minirender::Renderer) ---- */
MRX_API void* mrx_renderer_new(void);
MRX_API void mrx_renderer_free(void* r);
MRX_API int mrx_renderer_set_scene(void* r, void* scene);
MRX_API int mrx_renderer_set_size(void* r, int w, int h);
MRX_API int mrx_renderer_set_projection(void* r, const float* m16);
MRX_API int mrx_renderer_set_view(void* r, const float* m16);
MRX_API int mrx_renderer_set_light(void* r, const float* v3, int point);
MRX_API int mrx_renderer_set_lighting(void* r, int on);
MRX_API int mrx_renderer_set_texturing(void* r, int on);
MRX_API int mrx_renderer_set_save_normals(void* r, int on);
MRX_API int mrx_renderer_set_background(void* r, const float* rgb);
MRX_API int mrx_renderer_clear(void* r);
MRX_API int mrx_renderer_render(void* r);
/* In-place edits of a mesh, as reference programs make them between frames: TriMesh::applyTransform(), and
 * vertex i moved by (dx,dy,dz) in the mesh's own array. */
MRX_API int mrx_mesh_apply_transform(void* scene, int node);
MRX_API int mrx_mesh_move_vertex(void* scene, int node, int i, float dx, float dy, float dz);
MRX_API int mrx_renderer_paint_mesh(void* r, void* scene, int node, const float* xf);
/* Renderer::paintTriangle: three vertices of 8 floats each (position, normal, uv), already in view space */
MRX_API int mrx_renderer_paint_triangle(void* r, const float* v24, int world);
MRX_API int mrx_renderer_set_material(void* r, void* scene, int material);
MRX_API int mrx_renderer_get_image(void* r, float* out /* h*w*3 */);
MRX_API int mrx_renderer_get_depth(void* r, float* out /* h*w */);
MRX_API int mrx_renderer_get_normals(void* r, float* out /* h*w*3 */);
MRX_API int mrx_renderer_get_range(void* r, float* out /* h*w*3 */);

/* ---- image I/O on the path ---- */
MRX_API int mrx_quantize_rgb8(const float* image, int w, int h, uint8_t* out);

/* ---- asl::Matrix4 / projection helpers (so clients get bit-identical matrices) ---- */
MRX_API void mrx_mat_translate(float* out, float x, float y, float z);
MRX_API void mrx_mat_scale(float* out, float x, float y, float z);
MRX_API void mrx_mat_rotate_x(float* out, float a);
MRX_API void mrx_mat_rotate_y(float* out, float a);
MRX_API void mrx_mat_rotate_z(float* out, float a);
MRX_API void mrx_mat_rotate_axis(float* out, float x, float y, float z, float angle);
MRX_API void mrx_mat_rotate_vec(float* out, float x, float y, float z);
MRX_API void mrx_mat_mul(float* out, const float* a, const float* b);
MRX_API void mrx_mat_inverse(float* out, const float* a);
/* kind 0: ortho(l,r,b,t,n,f) 1: perspective(l,r,b,t,n,f) 2: frustum(fov,aspect,n,f) 3: frustumH 4: ortho(fov,aspect,n,f) */
MRX_API int mrx_projection(float* out, int kind, const float* args);
MRX_API void mrx_projection_cv(float* out, const float* K16, float w, float h, float n, float f);

/* ---- [product only] ---- */
MRX_API int mrx_renderer_set_device(void* r, int device);
MRX_API int mrx_renderer_set_row_range(void* r, int begin, int end);
MRX_API int mrx_renderer_invalidate_geometry(void* r);
MRX_API int mrx_renderer_prepare(void* r);               /* host part of render() only, no GPU */
MRX_API const void* mrx_renderer_scene_desc(void* r);    /* const mr_scene_desc*, valid until next prepare/render */
MRX_API const void* mrx_renderer_frame_desc(void* r);    /* const mr_frame* */
MRX_API void* mrx_renderer_context(void* r);             /* mr_ctx* (creates it) */
MRX_API int mrx_renderer_get_rgb8(void* r, uint8_t* out);
/* getImage()/getDepth() without the extra copy: pointer to the Renderer's own host mirror (valid until the next call) */
MRX_API const float* mrx_renderer_image_ptr(void* r);
MRX_API const float* mrx_renderer_depth_ptr(void* r);
MRX_API int mrx_renderer_synchronize(void* r);
MRX_API int mrx_save_ppm(const float* image, int w, int h, const char* filename);
/* mesh files: loadMesh() under `parent`; the loaded subtree's nodes get consecutive ids
 * in pre-order, the first of which is returned */
MRX_API int mrx_scene_load(void* scene, int parent, const char* filename);
MRX_API int mrx_scene_node_count(void* scene);
MRX_API int mrx_node_info(void* scene, int node, int32_t* is_mesh, int32_t* n_children, float* transform16);
MRX_API int mrx_mesh_material(void* scene, int node, float* out11, int32_t* tex_rows, int32_t* tex_cols);
MRX_API int mrx_save_stl(void* scene, int node, const char* filename);
MRX_API int mrx_save_xyz(const float* points, int w, int h, const float* m16, const char* filename);
MRX_API int mrx_triangulate(const int32_t* in, int n, int32_t* out);
MRX_API int mrx_load_ppm(const char* filename, float* out, int* rows, int* cols); /* out may be NULL to query the size */

#ifdef __cplusplus
}
#endif
#endif
