/*
 * raster_oracle.c — CPU restatement of minirender's render path. TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker the CUDA path is compared against. It is never linked into, loaded
 * by, or called from the product library (libminirender_b200.so); only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it.
 *
 * It consumes exactly the inputs the CUDA path consumes (mr_scene_desc + mr_frame from
 * include/minirender_b200.h) and restates, in plain scalar C with one IEEE binary32 rounding
 * per operator and the reference's left-to-right association, these reference functions:
 *
 *   Renderer::clear          src/Renderer.cpp:113-119
 *   Renderer::paintMesh      src/Renderer.cpp:334-381   (loops A, B, C)
 *   Renderer::paintTriangle  src/Renderer.cpp:163-309   (setup, raster loops D/E, shading)
 *   Renderer::clipTriangle   src/Renderer.cpp:131-161,  clip  :121-129
 *   htransform               src/Renderer.cpp:13-20
 *   Renderer::getRangeImage  src/Renderer.cpp:388-415
 *   savePPM quantiser        src/io.cpp:358-361
 *
 * Vector arithmetic follows ASL 1.11.14 as restated by third_party/asl_shim (Vec3/float is a
 * multiply by the reciprocal; min/max/clamp are the ternary forms; `pow` and `floor` resolve to
 * the double overloads in the reference translation unit, checked with nm on its object file).
 *
 * Pinning: the reference ships no tests or golden vectors, so this restatement is
 * pinned against the reference's own sources compiled unchanged (oracle/_ref, built by
 * oracle/Makefile from /root/reference) — bit-identical depth, float RGB and normals on every
 * scene of tests/scenes.py — and against the fixtures in tests/golden/ generated from that
 * build (tests/golden/make_golden.py). Build with -ffp-contract=off and without -march.
 */
#include "../include/minirender_b200.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;
typedef struct { v3 pos; v3 nrm; v2 uv; } vert;

typedef struct
{
	const mr_frame* f;
	int w, h;
	int persp;
	float* image;   /* h*w*3 */
	float* depth;   /* h*w */
	float* normals; /* h*w*3 or NULL */
	int32_t* winner; /* h*w submission ids or NULL */
	int row_begin, row_end;
	/* current material snapshot (Renderer.cpp:226-234) */
	const mr_material* mat;
	const mr_texture_desc* tex; /* NULL if the material has no texture */
	/* counters */
	int64_t records, clipped_in, frag_inside, frag_pass, bbox_px;
	int64_t cur_id; /* submission id of the triangle being painted */
} ostate;

/* ---- asl-style helpers (ternary forms decide NaN behaviour) ---- */
static inline float fmin_t(float a, float b) { return (a < b) ? a : b; }
static inline float fmax_t(float a, float b) { return (a > b) ? a : b; }
static inline float fclamp_t(float x, float a, float b) { return (x < a) ? a : (x > b) ? b : x; }

static inline v3 v3_make(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float v3_len(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline v3 v3_normalized(v3 a) { float q = 1 / v3_len(a); return v3_make(a.x * q, a.y * q, a.z * q); }
static inline v2 v2_make(float x, float y) { v2 r; r.x = x; r.y = y; return r; }

/* asl::Matrix4 * Vec3 on the top three rows of a row-major 3x4 (Renderer.cpp:345,348) */
static inline v3 affine(const float* m, v3 p)
{
	return v3_make(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3],
	               m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
	               m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}

/* Renderer.cpp:13-20 */
static inline v3 htransform(const float* m, v3 p)
{
	float iw = 1 / (m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15]);
	return v3_make((m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3]) * iw,
	               (m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7]) * iw,
	               (m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]) * iw);
}

static void paint_triangle(ostate* s, const vert* v0, const vert* v1, const vert* v2_, int world);

/* Renderer.cpp:121-129 */
static vert clip_edge(float z, const vert* a, const vert* b)
{
	float k = (fabs(b->pos.z - a->pos.z) < 1e-6f) ? 0.5f : (z - a->pos.z) / (b->pos.z - a->pos.z);
	float k1 = 1 - k;
	vert v;
	v.pos = v3_add(v3_scale(b->pos, k), v3_scale(a->pos, k1));
	v.nrm = v3_add(v3_scale(b->nrm, k), v3_scale(a->nrm, k1));
	v.uv = v2_make(b->uv.x * k + a->uv.x * k1, b->uv.y * k + a->uv.y * k1);
	return v;
}

/* Renderer.cpp:131-161 */
static void clip_triangle(ostate* s, float z, vert v[3])
{
	if (v[0].pos.z > z && v[1].pos.z > z && v[2].pos.z > z)
		return;
	while (v[0].pos.z < v[1].pos.z || v[0].pos.z < v[2].pos.z)
	{
		vert t = v[0]; v[0] = v[1]; v[1] = t;
		t = v[0]; v[0] = v[2]; v[2] = t;
	}
	if (v[1].pos.z > z)
	{
		vert v02 = clip_edge(z, &v[0], &v[2]);
		vert v12 = clip_edge(z, &v[1], &v[2]);
		paint_triangle(s, &v02, &v12, &v[2], 0);
	}
	else if (v[2].pos.z > z)
	{
		vert v01 = clip_edge(z, &v[0], &v[1]);
		vert v12 = clip_edge(z, &v[1], &v[2]);
		paint_triangle(s, &v01, &v[1], &v12, 0);
	}
	else
	{
		vert v01 = clip_edge(z, &v[0], &v[1]);
		vert v02 = clip_edge(z, &v[0], &v[2]);
		paint_triangle(s, &v01, &v[1], &v[2], 0);
		s->cur_id++; /* second sub-triangle gets the next submission id */
		paint_triangle(s, &v01, &v[2], &v02, 0);
		s->cur_id--;
	}
}

/* Renderer.cpp:163-309 */
static void paint_triangle(ostate* s, const vert* v0, const vert* v1, const vert* v2_, int world)
{
	const mr_frame* f = s->f;
	v3 vertices[3] = { v0->pos, v1->pos, v2_->pos };
	v3 nrm[3] = { v0->nrm, v1->nrm, v2_->nrm };
	v2 tc[3] = { v0->uv, v1->uv, v2_->uv };
	float znear = f->znear;

	if (world && (vertices[0].z > znear || vertices[1].z > znear || vertices[2].z > znear))
	{
		vert verts[3];
		if (vertices[0].z > znear && vertices[1].z > znear && vertices[2].z > znear)
			return;
		verts[0] = *v0; verts[1] = *v1; verts[2] = *v2_;
		s->clipped_in++;
		clip_triangle(s, znear, verts);
		return;
	}

	float w = (float)s->w, h = (float)s->h;
	v3 ndc[3];
	v2 p[3];
	v2 pmin = v2_make(1e30f, 1e30f), pmax = v2_make(-1e30f, -1e30f);
	int i;
	for (i = 0; i < 3; i++)
		ndc[i] = htransform(f->projection, vertices[i]);
	for (i = 0; i < 3; i++)
	{
		p[i].x = (1 + ndc[i].x) * (w / 2);
		p[i].y = (1 - ndc[i].y) * (h / 2);
		pmin = v2_make(fmin_t(pmin.x, p[i].x), fmin_t(pmin.y, p[i].y));
		pmax = v2_make(fmax_t(pmax.x, p[i].x), fmax_t(pmax.y, p[i].y));
	}
	if (pmax.x < 0 || pmax.y < 0 || pmin.x > w || pmin.y > h)
		return;

	/* (p0-p1) ^ (p2-p1) */
	float a = (p[0].x - p[1].x) * (p[2].y - p[1].y) - (p[0].y - p[1].y) * (p[2].x - p[1].x);
	if (a <= 0)
		return;
	float i2a = (a == 0) ? 0.0f : -1.0f / a;
	/* perpend(v) = (-v.y, v.x), then * i2a */
	v2 n1 = v2_make(-(p[0].y - p[2].y) * i2a, (p[0].x - p[2].x) * i2a);
	v2 n2 = v2_make(-(p[1].y - p[0].y) * i2a, (p[1].x - p[0].x) * i2a);

	pmin.x = fclamp_t(pmin.x, 0.f, w - 1);
	pmax.x = fclamp_t(pmax.x, 0.f, w - 1);
	pmin.y = fclamp_t(pmin.y, 0.f, h - 1);
	pmax.y = fclamp_t(pmax.y, 0.f, h - 1);

	float zz[4] = { ndc[0].z, ndc[1].z, ndc[2].z, 1 };
	float iz[4] = { -1 / vertices[0].z, -1 / vertices[1].z, -1 / vertices[2].z, 1 };

	int persp = s->persp;
	const mr_material* m = s->mat;
	int hasspecular = m->shininess != 0;
	int hastexture = f->texturing && s->tex && s->tex->rows > 0;
	v3 color = v3_make(m->diffuse[0], m->diffuse[1], m->diffuse[2]);
	v3 emissive = v3_make(m->emissive[0], m->emissive[1], m->emissive[2]);
	v3 mspecular = v3_make(m->specular[0], m->specular[1], m->specular[2]);
	float shininess = m->shininess;
	v3 light = v3_make(f->light[0], f->light[1], f->light[2]);
	float k[4] = { 0, 0, 0, 0 };
	float y, x;

	s->records++;

	for (y = (float)(floor(pmin.y) + 0.5f); y <= pmax.y + 0.5f; y++)
	{
		float ptx = (float)(floor(pmin.x) + 0.5f);
		float e1 = n1.x * (ptx - p[2].x) + n1.y * (y - p[2].y);
		float e2 = n2.x * (ptx - p[0].x) + n2.y * (y - p[0].y);
		for (x = ptx; x <= pmax.x + 0.5f; x++, e1 += n1.x, e2 += n2.x)
		{
			s->bbox_px++;
			if (e1 < 0 || e2 < 0 || 1 - e1 - e2 < 0)
				continue;
			s->frag_inside++;
			k[0] = 1.f - e1 - e2;
			k[1] = e1;
			k[2] = e2;
			float z;
			if (persp)
			{
				z = 1.f / (k[0] * iz[0] + k[1] * iz[1] + k[2] * iz[2]);
				k[0] *= iz[0] * z;
				k[1] *= iz[1] * z;
				k[2] *= iz[2] * z;
			}
			else
				z = k[0] * zz[0] + k[1] * zz[1] + k[2] * zz[2] + k[3] * zz[3];

			int pi = (int)y, pj = (int)x;
			if (pi < s->row_begin || pi >= s->row_end)
				continue; /* strip rendering: rows outside the strip belong to another rank */
			float* pixdepth = &s->depth[(size_t)pi * s->w + pj];
			if (z < *pixdepth)
			{
				*pixdepth = z;
				s->frag_pass++;
				if (s->winner)
					s->winner[(size_t)pi * s->w + pj] = (int32_t)s->cur_id;
				if (hastexture)
				{
					v2 uv = v2_make(tc[0].x * k[0] + tc[1].x * k[1] + tc[2].x * k[2],
					                tc[0].y * k[0] + tc[1].y * k[1] + tc[2].y * k[2]);
					float fy = uv.y - (float)floor(uv.y), fx = uv.x - (float)floor(uv.x);
					int ti = (int)(fy * s->tex->rows), tj = (int)(fx * s->tex->cols);
					/* fract() can return exactly 1.0f for tiny negative inputs, which makes the
					   reference index one past the end (UB). Both this oracle and the CUDA path
					   clamp instead (documented divergence). */
					if (ti > s->tex->rows - 1) ti = s->tex->rows - 1;
					if (tj > s->tex->cols - 1) tj = s->tex->cols - 1;
					if (ti < 0) ti = 0;
					if (tj < 0) tj = 0;
					const float* t = s->tex->texels + 3 * ((size_t)ti * s->tex->cols + tj);
					color = v3_make(t[0], t[1], t[2]);
				}
				v3 value = emissive;
				if (f->lighting)
				{
					v3 position = v3_add(v3_add(v3_scale(vertices[0], k[0]), v3_scale(vertices[1], k[1])), v3_scale(vertices[2], k[2]));
					v3 lightdir = f->light_is_point ? v3_normalized(v3_sub(light, position)) : light;
					v3 normal = v3_add(v3_add(v3_scale(nrm[0], k[0]), v3_scale(nrm[1], k[1])), v3_scale(nrm[2], k[2]));
					float nl = v3_dot(normal, lightdir);
					float d = ((0.0f > nl) ? 0.0f : nl) / v3_len(normal) + f->ambient;
					value = v3_add(value, v3_scale(color, d));
					if (hasspecular)
					{
						v3 viewdir = v3_normalized(position);
						v3 hv = v3_sub(lightdir, viewdir);
						float hn = v3_dot(hv, normal);
						float base = ((hn > 0.0f) ? hn : 0.0f) / (v3_len(hv) * v3_len(normal));
						float specular = (float)pow((double)base, (double)shininess);
						value = v3_add(value, v3_scale(mspecular, specular));
					}
					if (f->save_normals && s->normals)
					{
						float* pn = &s->normals[3 * ((size_t)pi * s->w + pj)];
						pn[0] = normal.x; pn[1] = normal.y; pn[2] = normal.z;
					}
				}
				float* px = &s->image[3 * ((size_t)pi * s->w + pj)];
				px[0] = value.x; px[1] = value.y; px[2] = value.z;
			}
		}
	}
}

/* Counters of the last oracle_render (work-count cross-checks against mr_stats). */
static int64_t g_counters[8];

MR_API void oracle_last_counters(int64_t out[8]) { memcpy(out, g_counters, sizeof(g_counters)); }

/*
 * Renders `frame` over `scene` into caller-owned buffers. `normals` and `winner` may be NULL.
 * winner[i] receives the submission id of the fragment that owns pixel i:
 * 2*(global triangle index over the renderable list) + (1 for the second triangle produced by a
 * near-plane clip), or -1 for background. With frame->keep the buffers are depth-tested against
 * and kept instead of being cleared.
 */
MR_API int oracle_render(const mr_scene_desc* scene, const mr_frame* frame, int w, int h,
                         float* image, float* depth, float* normals, int32_t* winner)
{
	ostate s;
	size_t npix = (size_t)w * h, i;
	int r;
	if (!scene || !frame || !image || !depth || w <= 0 || h <= 0)
		return MR_E_INVALID;
	memset(&s, 0, sizeof(s));
	s.f = frame;
	s.w = w; s.h = h;
	s.image = image; s.depth = depth; s.normals = normals; s.winner = winner;
	s.persp = frame->projection[15] == 0;
	s.row_begin = 0; s.row_end = h;
	if (frame->row_end > frame->row_begin)
	{
		s.row_begin = frame->row_begin < 0 ? 0 : frame->row_begin;
		s.row_end = frame->row_end > h ? h : frame->row_end;
	}
	if (!frame->keep) /* Renderer::clear, Renderer.cpp:113-119 (restricted to the strip) */
	{
		for (i = (size_t)s.row_begin * w; i < (size_t)s.row_end * w; i++)
		{
			image[3 * i] = frame->background[0];
			image[3 * i + 1] = frame->background[1];
			image[3 * i + 2] = frame->background[2];
			depth[i] = 1e11f;
			if (winner) winner[i] = -1;
			if (normals && frame->save_normals)
			{
				normals[3 * i] = 0; normals[3 * i + 1] = 0; normals[3 * i + 2] = 1;
			}
		}
	}
	(void)npix;

	int64_t tri_base = 0;
	for (r = 0; r < frame->n_renderables; r++)
	{
		const mr_renderable* rd = &frame->renderables[r];
		if (rd->mesh < 0 || rd->mesh >= scene->n_meshes || rd->material < 0 || rd->material >= frame->n_materials)
			return MR_E_INVALID;
		const mr_mesh_desc* m = &scene->meshes[rd->mesh];
		s.mat = &frame->materials[rd->material];
		s.tex = (s.mat->texture >= 0 && s.mat->texture < scene->n_textures) ? &scene->textures[s.mat->texture] : NULL;
		/* loops A and B, Renderer.cpp:344-348 */
		v3* tv = (v3*)malloc(sizeof(v3) * (size_t)(m->n_positions > 0 ? m->n_positions : 1));
		v3* tn = (v3*)malloc(sizeof(v3) * (size_t)(m->n_normals > 0 ? m->n_normals : 1));
		int j, t;
		if (!tv || !tn) { free(tv); free(tn); return MR_E_NOMEM; }
		for (j = 0; j < m->n_positions; j++)
			tv[j] = affine(rd->modelview, v3_make(m->positions[3 * j], m->positions[3 * j + 1], m->positions[3 * j + 2]));
		for (j = 0; j < m->n_normals; j++)
			tn[j] = affine(rd->normalmat, v3_make(m->normals[3 * j], m->normals[3 * j + 1], m->normals[3 * j + 2]));
		int hasuv = m->n_texcoords > 0 && m->idx_uv != NULL;
		/* loop C, Renderer.cpp:351-380 */
		for (t = 0; t < m->n_triangles; t++)
		{
			vert q[3];
			int c;
			for (c = 0; c < 3; c++)
			{
				q[c].pos = tv[m->idx_pos[3 * t + c]];
				q[c].nrm = tn[m->idx_nrm[3 * t + c]];
				if (hasuv)
				{
					int ui = m->idx_uv[3 * t + c];
					q[c].uv = v2_make(m->texcoords[2 * ui], m->texcoords[2 * ui + 1]);
				}
				else
					q[c].uv = v2_make(0, 0);
			}
			s.cur_id = 2 * (tri_base + t);
			paint_triangle(&s, &q[0], &q[1], &q[2], 1);
		}
		tri_base += m->n_triangles;
		free(tv);
		free(tn);
	}
	g_counters[0] = tri_base;
	g_counters[1] = s.records;
	g_counters[2] = s.clipped_in;
	g_counters[3] = s.bbox_px;
	g_counters[4] = s.frag_inside;
	g_counters[5] = s.frag_pass;
	return MR_OK;
}

/* Renderer::getRangeImage, Renderer.cpp:388-415 */
MR_API int oracle_range_image(const float* P, float znear, const float* depth, int w, int h, float* xyz)
{
	int persp = P[15] == 0;
	float zfar = persp ? P[11] / (P[10] + 1) : (P[11] - 1) / P[10];
	float fardepth = persp ? zfar : 1.0f;
	float fw = (float)w, fh = (float)h;
	int i, j;
	(void)znear;
	for (i = 0; i < h; i++)
		for (j = 0; j < w; j++)
		{
			float d = depth[(size_t)i * w + j];
			float* o = &xyz[3 * ((size_t)i * w + j)];
			if (d > fardepth)
			{
				o[0] = 0; o[1] = 0; o[2] = 0;
			}
			else
			{
				float u = (j + 0.5f) / (fw / 2) - 1;
				float v = -(i + 0.5f) / (fh / 2) + 1;
				float z = -d;
				o[0] = -(u + P[2]) * z / P[0];
				o[1] = -(v + P[6]) * z / P[5];
				o[2] = z;
			}
		}
	return MR_OK;
}

/* savePPM's quantiser, io.cpp:358-361: (byte)clamp(v*255, 0, 255), truncation */
MR_API int oracle_quantize_rgb8(const float* image, size_t n_floats, uint8_t* out)
{
	size_t i;
	for (i = 0; i < n_floats; i++)
	{
		float v = image[i] * 255.0f;
		v = (v < 0.0f) ? 0.0f : (v > 255.0f) ? 255.0f : v;
		out[i] = (uint8_t)v;
	}
	return MR_OK;
}
