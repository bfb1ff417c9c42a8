/*
 * minirender_b200.h — C ABI of the B200 rasterization path.
 *
 * The reference (aslze/minirender) has no FFI: its only seam is the C++ class
 * minirender::Renderer (reference include/minirender/Renderer.h:17-64). Our drop-in
 * Renderer (include/minirender/Renderer.h in this repo) keeps that class API and calls the
 * functions below from inside render()/getImage()/getDepth(); nothing else in the product
 * touches CUDA. Everything here is plain C: pointers, sizes, ints, floats. No torch, no ASL.
 *
 * What each entry point replaces in the reference:
 *   mr_set_size        Renderer::setSize            src/Renderer.cpp:100-106
 *   mr_upload_scene    the arrays paintMesh reads   src/Renderer.cpp:341-380, Scene.h:73-87
 *   mr_render          clear + loops A..E           src/Renderer.cpp:113-119, 163-309, 344-380
 *   mr_read_image      Renderer::getImage           src/Renderer.cpp:383-386
 *   mr_read_depth      Renderer::getDepth           include/minirender/Renderer.h:60
 *   mr_read_normals    Renderer::getNormalsImage    include/minirender/Renderer.h:63
 *   mr_read_range      Renderer::getRangeImage      src/Renderer.cpp:388-415
 *   mr_read_rgb8       savePPM's quantiser          src/io.cpp:358-361
 *
 * Conventions: every function returns 0 on success or a negative MR_E_* code; the message
 * for the last failure on a context is mr_last_error(ctx). No exceptions cross this ABI.
 * A context owns all of its device memory, is bound to one GPU and one CUDA stream, and is
 * not thread-safe (one context per host thread, like one reference Renderer per thread).
 * Host pointers in descriptors are borrowed for the duration of the call only.
 * mr_render is asynchronous (stream ordered); the mr_read_* calls synchronise.
 * There is no CPU fallback: without a usable CUDA device mr_create fails.
 */
#ifndef MINIRENDER_B200_H
#define MINIRENDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MR_API __attribute__((visibility("default")))
#else
#define MR_API
#endif

#define MR_ABI_VERSION 2 /* 2: mr_stats grew (chk_entries, chk_demand, d2h_bytes); mr_download, mr_read_image_dirty_*, mr_ipc_export_slot, mr_clear_rows_slot, mr_output_slot */

enum
{
	MR_OK = 0,
	MR_E_INVALID = -1,   /* bad argument / descriptor */
	MR_E_CUDA = -2,      /* CUDA runtime error (see mr_last_error) */
	MR_E_NO_DEVICE = -3, /* no CUDA device / driver: the product does not fall back to the CPU */
	MR_E_NO_SCENE = -4,  /* mr_render before mr_upload_scene */
	MR_E_OVERFLOW = -5,  /* internal queue overflow that could not be recovered by regrowing */
	MR_E_NOMEM = -6
};

typedef struct mr_ctx mr_ctx;

/* One TriMesh's geometry (reference Scene.h:73-87). Index arrays hold 3 ints per triangle
 * and index into this mesh's own attribute arrays. idx_nrm is required (the reference reads
 * normalsI unconditionally, Renderer.cpp:367-369). The mesh is textured-capable only when
 * both n_texcoords > 0 and idx_uv != NULL (Renderer.cpp:371). */
typedef struct mr_mesh_desc
{
	const float* positions; /* 3*n_positions floats, xyz packed (asl::Vec3 layout) */
	const float* normals;   /* 3*n_normals floats */
	const float* texcoords; /* 2*n_texcoords floats, or NULL */
	const int32_t* idx_pos; /* 3*n_triangles */
	const int32_t* idx_nrm; /* 3*n_triangles */
	const int32_t* idx_uv;  /* 3*n_triangles, or NULL */
	int32_t n_positions;
	int32_t n_normals;
	int32_t n_texcoords;
	int32_t n_triangles;
} mr_mesh_desc;

/* Float RGB texture, row-major rows x cols x 3 (asl::Array2<asl::Vec3>, Scene.h:42). */
typedef struct mr_texture_desc
{
	const float* texels;
	int32_t rows;
	int32_t cols;
} mr_texture_desc;

typedef struct mr_scene_desc
{
	const mr_mesh_desc* meshes;
	const mr_texture_desc* textures;
	int32_t n_meshes;
	int32_t n_textures;
} mr_scene_desc;

/* Material snapshot (reference Scene.h:38-45; opacity is never read by the renderer). */
typedef struct mr_material
{
	float diffuse[3];
	float specular[3];
	float emissive[3];
	float shininess;
	int32_t texture; /* index into mr_scene_desc.textures, or -1 */
	int32_t _pad;
} mr_material;

/* One entry of the flattened scene (reference Scene.h:47-53), in submission order. The two
 * matrices are computed on the host with the reference's own expressions
 * (Renderer.cpp:337-338): modelview = view * world, normalmat = modelview.inverse().t();
 * only their top three rows are used (affine asl::Matrix4 * Vec3). Row-major 3x4. */
typedef struct mr_renderable
{
	float modelview[12];
	float normalmat[12];
	int32_t mesh;     /* index into mr_scene_desc.meshes */
	int32_t material; /* index into mr_frame.materials */
} mr_renderable;

/* Per-frame state: everything Renderer::render() derives before its renderable loop
 * (Renderer.cpp:313-326) plus the renderer's switches (Renderer.h:48-55). */
typedef struct mr_frame
{
	float projection[16];  /* row-major 4x4 */
	const mr_renderable* renderables;
	const mr_material* materials;
	int32_t n_renderables;
	int32_t n_materials;
	float light[3];        /* _lightdir: view-space position (point) or normalised direction */
	int32_t light_is_point;
	float ambient;         /* Scene::ambientLight */
	float znear;           /* Renderer.cpp:325-326 */
	int32_t lighting;
	int32_t texturing;
	int32_t save_normals;
	float background[3];
	int32_t row_begin;     /* render only image rows [row_begin,row_end); 0,0 = whole image.  */
	int32_t row_end;       /* Used for strip sharding across GPUs; other rows are untouched.   */
	int32_t keep;          /* 0: clear first (render()); 1: depth-test against and keep the
	                          current buffers (immediate-mode paintMesh after clear()) */
} mr_frame;

/* Counters of the last completed mr_render on this context (valid after a sync/read). */
typedef struct mr_stats
{
	int64_t triangles_in;      /* triangles submitted */
	int64_t records;           /* set-up triangles that survived near test, reject, cull and the zero-coverage test */
	int64_t clipped_in;        /* input triangles that crossed the near plane */
	int64_t bin_entries;       /* (tile, triangle) pairs */
	int64_t zero_coverage;     /* set-up triangles dropped because they cover no pixel centre */
	int32_t tiles_x, tiles_y;
	int32_t regrows;           /* times a queue had to be regrown and the frame re-run */
	int32_t kernels_launched;  /* kernel launches issued for the frame */
	float   ms_kernel[8];      /* per-stage device time of the last mr_profile_frame:
	                              1 geometry (k_geom), 4 tile resolve + shade (k_raster), 5 whole frame; others 0 */
	int64_t h2d_bytes;         /* host->device bytes the last mr_render copied (per-frame tables) */
	int64_t clusters;          /* 32-triangle clusters of the frame ... */
	int64_t clusters_visible;  /* ... and how many of them survived cluster culling */
	int64_t tiles_stored;      /* 16x16 tiles the tile kernel wrote (with sparse remote stores: the touched ones only) */
	int64_t chk_entries;       /* edge-chain checkpoints written for wide triangles (0: the frame ran without k_chain) */
	int64_t chk_demand;        /* ... and how many the frame's wide triangles asked for */
	int64_t d2h_bytes;         /* bytes the last mr_read_image_begin / mr_read_image_dirty_begin copied to the host */
} mr_stats;

MR_API int mr_abi_version(void);
MR_API int mr_device_count(void);

MR_API mr_ctx* mr_create(int device, int* status);
MR_API void mr_destroy(mr_ctx* ctx);
MR_API const char* mr_last_error(const mr_ctx* ctx);

/* Use an externally owned CUDA stream (cudaStream_t / CUstream as void*); NULL = ctx-owned. */
MR_API int mr_set_stream(mr_ctx* ctx, void* cuda_stream);
MR_API int mr_set_size(mr_ctx* ctx, int width, int height);
MR_API int mr_upload_scene(mr_ctx* ctx, const mr_scene_desc* scene);

MR_API int mr_render(mr_ctx* ctx, const mr_frame* frame);
/* n frames over the same scene. `sink` (may be NULL) is called once per frame, in order, with device
 * pointers to the finished frame's float image / depth; they stay valid until the sink returns. With two output sets
 * (mr_set_output_slots(ctx, 2)) frame i + 1 is already running on the device while the sink sees frame i - no
 * device-wide synchronisation per frame; with one set every frame is finished and delivered before the next is launched.
 * A frame that overflowed its spill list is rendered again before it is delivered. Without a sink the frames are
 * launched back to back and the call returns when the last one is complete. */
typedef void (*mr_frame_sink)(void* user, int index, const float* d_image, const float* d_depth);
MR_API int mr_render_batch(mr_ctx* ctx, int n, const mr_frame* frames, mr_frame_sink sink, void* user);

MR_API int mr_synchronize(mr_ctx* ctx);
/* blocking copy out of a device pointer the library handed out (a sink's d_image / d_depth, mr_device_buffers) */
MR_API int mr_download(mr_ctx* ctx, void* host, const void* d_src, size_t bytes);

/* Blocking device->host reads of the last frame (full image, row-major, row 0 = top). */
MR_API int mr_read_image(mr_ctx* ctx, float* host_rgb /* h*w*3 */);
MR_API int mr_read_depth(mr_ctx* ctx, float* host_depth /* h*w */);
MR_API int mr_read_normals(mr_ctx* ctx, float* host_nrm /* h*w*3 */);
MR_API int mr_read_range(mr_ctx* ctx, const float* projection16, float znear, float* host_xyz /* h*w*3 */);
MR_API int mr_read_rgb8(mr_ctx* ctx, uint8_t* host_rgb8 /* h*w*3 */);
/* Same reads into caller-provided *pinned or pageable* memory without the final sync. */
MR_API int mr_read_image_async(mr_ctx* ctx, float* host_rgb);
MR_API int mr_read_rows_async(mr_ctx* ctx, float* host_rgb, float* host_depth, int row_begin, int row_end);

/* Overlapping the host copy of frame i with the kernels of frame i+1 (a turntable / view batch that
 * needs every image on the host): with two output slots, successive mr_render calls alternate
 * between two sets of image/depth buffers. mr_read_image_begin starts the device->host copy of the
 * newest frame on a separate copy stream (ordered after that frame, `host_rgb` should be
 * page-locked) and returns a ticket; mr_read_wait blocks until that copy has landed. A slot is not
 * rendered into again before its pending copy has finished. */
MR_API int mr_set_output_slots(mr_ctx* ctx, int n /* 1 or 2 */);
MR_API int mr_read_image_begin(mr_ctx* ctx, float* host_rgb /* h*w*3 */, int* ticket);
MR_API int mr_read_wait(mr_ctx* ctx, int ticket);
/* The same for a host buffer the application keeps between frames (a turntable writing every image into the same one or
 * two page-locked buffers): the library remembers which rectangle of `host_rgb` is not background and copies only the
 * union of that and the rectangle the new frame may have drawn into (known from the frame's counters: k_geom tracks the
 * tiles of every triangle that emits a fragment or is binned); everything outside already holds the frame's background
 * on both sides. The buffer ends up bit-identical to a full mr_read_image. The first use of a buffer, a changed size or
 * background, kept (immediate-mode) or strip frames copy the whole image. Waits for the frame's kernels (its counters)
 * before it returns; the copy itself is asynchronous like mr_read_image_begin's. mr_stats.d2h_bytes = bytes copied. */
MR_API int mr_read_image_dirty_begin(mr_ctx* ctx, float* host_rgb /* h*w*3, persistent */, int* ticket);
/* The library assumes that nobody else writes into such a buffer between two calls. After writing into it, or after
 * freeing it (another allocation may get the same address), make the library forget it (NULL: all buffers): the next
 * mr_read_image_dirty_begin into it copies the whole image again. */
MR_API int mr_read_image_dirty_forget(mr_ctx* ctx, const float* host_rgb);

/* Device pointers of the current output buffers (for interop: NCCL gather, checksums). */
MR_API int mr_device_buffers(mr_ctx* ctx, void** d_image, void** d_depth, void** d_normals);
/* Overwrite rows [row_begin,row_end) of the output buffers from device memory (strip gather). */
MR_API int mr_write_rows(mr_ctx* ctx, const void* d_image_rows, const void* d_depth_rows, int row_begin, int row_end);
/* Let tile stores of rows [row_begin,row_end) land directly in a peer GPU's framebuffer
 * (pointers obtained from that peer's mr_device_buffers through CUDA IPC / symmetric memory). */
MR_API int mr_set_remote_target(mr_ctx* ctx, void* d_peer_image, void* d_peer_depth);
/* CUDA IPC plumbing for that: export this context's image+depth buffers as two 64-byte
 * cudaIpcMemHandle_t (128 bytes), open a peer's handles on this context's device, close them. */
MR_API int mr_ipc_export(mr_ctx* ctx, void* handles128);
MR_API int mr_ipc_open(mr_ctx* ctx, const void* handles128, void** d_image, void** d_depth);
/* With two output slots: the handles of slot 0 or 1, and the slot the newest frame was rendered into. A gathering rank
 * that alternates between two framebuffers lets its peers store frame i + 1 while it still consumes and clears frame i. */
MR_API int mr_ipc_export_slot(mr_ctx* ctx, int slot, void* handles128);
MR_API int mr_output_slot(mr_ctx* ctx);
MR_API int mr_ipc_close(mr_ctx* ctx, void* d_image, void* d_depth);

/* Joining the ranks of a strip-sharded frame on the device, without a collective and without the host:
 * mr_sync_words gives a small zeroed device buffer of 32-bit words on this context's GPU (exported to the peers with
 * mr_ipc_export_ptr / opened with mr_ipc_open_ptr). mr_stream_signal makes the context's stream store `value` into a
 * word - local or a peer's - once everything enqueued before it (a rank's tile stores into the gathering rank's
 * framebuffer, say) is visible system-wide; mr_stream_wait makes the stream wait until each of `n` words has reached
 * `value`. A frame of the strip mode is then: every rank renders its rows into rank 0's framebuffer and signals
 * arrived[rank] = frame + 1; rank 0 waits for all of them; before the next frame the peers wait for rank 0's go word. */
MR_API int mr_sync_words(mr_ctx* ctx, int n_words, void** d_words);
MR_API int mr_ipc_export_ptr(mr_ctx* ctx, const void* d_ptr, void* handle64);
MR_API int mr_ipc_open_ptr(mr_ctx* ctx, const void* handle64, void** d_ptr);
MR_API int mr_ipc_close_ptr(mr_ctx* ctx, void* d_ptr);
MR_API int mr_stream_signal(mr_ctx* ctx, void* d_word, uint32_t value);
/* One shot: the next mr_render waits for *d_word >= value between its geometry kernel and its tile kernel, so that a
 * peer's geometry for frame i + 1 runs while the gathering rank still owns the framebuffer of frame i. */
MR_API int mr_set_raster_gate(mr_ctx* ctx, const void* d_word, uint32_t value);
/* With a remote target set: tiles nothing was drawn into are not stored (flag != 0); the owner of the framebuffer
 * writes the clear values of those rows itself (mr_clear_rows) before it lets the peers in. */
MR_API int mr_set_sparse_remote_stores(mr_ctx* ctx, int flag);
/* Clear values (background / 1e11) into rows [row_begin,row_end) of the current output buffers, stream ordered. */
MR_API int mr_clear_rows(mr_ctx* ctx, const float* background3, int row_begin, int row_end);
MR_API int mr_clear_rows_slot(mr_ctx* ctx, int slot, const float* background3, int row_begin, int row_end);
MR_API int mr_stream_wait(mr_ctx* ctx, const void* d_words, int n, uint32_t value);

/* Page-lock caller memory so the mr_read_* copies run at full PCIe rate (optional). */
MR_API int mr_host_register(void* host, size_t bytes);
MR_API int mr_host_unregister(void* host);

/* Verification aid: with flags & 1, mr_render also records for every pixel the submission id of
 * the triangle that owns it (2 * triangle instance index + clip sub-triangle, -1 = background),
 * which is what equal-depth ties are resolved on. Read it back with mr_read_winner_ids.
 * flags & 4 turns cluster culling off, flags & 8 the standard-perspective vertex path and flags & 16 the tight scan of
 * small triangles (every cluster is set up / every corner goes through the general htransform / every pixel centre of
 * the reference's loops is tested): the image must not change, which is what the tests use them for. (The environment
 * variables MR_NO_CLUSTER_CULL / MR_NO_STD_PROJ / MR_NO_TIGHT_SCAN do the same for a whole process.)
 * flags & 64: no warp starts on a cluster of its own cull round (every cluster goes through the work list).
 * flags & 128 / 256: edge-chain checkpoints of wide triangles (k_chain) forced on from the first frame / off (normally a
 * frame runs k_chain when the frame before it had wide triangles; MR_NO_CHAIN_CHECKPOINTS in the environment = 256).
 * flags & 512: every record is written with its fifth pair (texture coordinates, rows, flags) whether or not anyone reads it.
 * flags & 2048: records keep the corner positions and shading interpolates them like the reference (Renderer.cpp:281) also
 * under the standard perspective, where a pixel's position is otherwise taken from its own ray and depth. */
MR_API int mr_set_debug(mr_ctx* ctx, int flags);
MR_API int mr_read_winner_ids(mr_ctx* ctx, int32_t* host_ids /* h*w */);

MR_API int mr_get_stats(mr_ctx* ctx, mr_stats* out);
/* Evict the L2 cache by writing a 256 MiB scratch buffer on the context's stream (benchmark
 * hygiene: cold-cache timing between steps). */
MR_API int mr_flush_l2(mr_ctx* ctx);
/* Benchmark aid: the next mr_render records these two CUDA events (cudaEvent_t) on the render stream
 * immediately before its first and after its last kernel launch, so that a device-time bracket contains the
 * frame's kernels and nothing of the caller's host work. One shot: cleared by that mr_render. */
MR_API int mr_set_timing_events(mr_ctx* ctx, void* cuda_event_start, void* cuda_event_stop);
/* Render `frame` `repeats` times with CUDA events between the stages; fills mr_stats.ms_kernel
 * with the averages. With mr_set_debug flag 2 the L2 is flushed before every repeat. */
MR_API int mr_profile_frame(mr_ctx* ctx, const mr_frame* frame, int repeats);

#ifdef __cplusplus
}
#endif
#endif
