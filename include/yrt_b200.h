/*
 * yrt_b200.h — C ABI of the B200-native render path (libyrt_b200.so).
 *
 * The reference (sebcossu/yocto_raytracing) has no plugin / FFI interface; its hot
 * path is the C++ call pair in main():
 *     build_bvh(scn, false);                                  src/raytrace.cpp:278  (decl src/scene.h:238)
 *     auto hdr = raytrace(scn, {amb,amb,amb}, resolution, samples);   src/raytrace.cpp:282  (def :213)
 * This header is what a maintainer binds instead of those two calls: the host keeps
 * load_scene() (src/scene.cpp:113) and save_hdr_or_ldr() (src/image.cpp:81), flattens
 * the loaded `scene` (src/scene.h:136-155) into the SoA `yrt_scene_desc` below and
 * calls yrt_scene_create() + yrt_render().  See INTEGRATION.md for the stub.
 *
 * Conventions
 *  - plain C, plain pointers and sizes; no CUDA / torch types in any signature
 *    (a CUDA stream crosses as `void*`, a device pointer as `void*`).
 *  - all `const T*` inputs are HOST memory borrowed for the duration of the call.
 *  - every function returning int returns YRT_OK (0) or a negative yrt_status;
 *    yrt_last_error() gives the message (thread-local).  The reference's own error
 *    convention is printf + exit(1) (src/scene.cpp:119-122); the host shim mirrors it.
 *  - there is NO CPU fallback: without a usable CUDA device every compute entry point
 *    fails with YRT_ERR_NO_DEVICE.
 */
#ifndef YRT_B200_H_
#define YRT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YRT_ABI_VERSION 3   /* 3: + yrt_frame_truncated_paths; scenes with non-rigid instance frames are rendered (ABI 2 refused them unless opted in) */

typedef enum yrt_status {
    YRT_OK = 0,
    YRT_ERR_INVALID = -1,      /* bad argument / malformed scene description        */
    YRT_ERR_NO_DEVICE = -2,    /* no CUDA device, or fewer than requested            */
    YRT_ERR_CUDA = -3,         /* a CUDA runtime call or kernel failed               */
    YRT_ERR_UNSUPPORTED = -4,  /* valid input the path does not cover (see message)  */
    YRT_ERR_OOM = -5
} yrt_status;

/* element kind of a shape: the reference keeps three index vectors per shape
 * (src/scene.h:36-38) and dispatches triangles > lines > points (src/scene.cpp:405-427);
 * every shape the loader produces holds exactly one kind, which is what we require. */
enum { YRT_TRIANGLES = 0, YRT_LINES = 1, YRT_POINTS = 2 };

/*
 * Flattened scene (replaces the pointer graph `scene`, src/scene.h:136-155).
 * SoA, int32 ids, float32 data.  Vertex indices in elem_idx are LOCAL to the shape
 * (0 .. shape_vert_cnt-1), exactly the values held in shape::triangles/lines/points.
 */
typedef struct yrt_scene_desc {
    int32_t n_shapes, n_instances, n_materials, n_textures;
    int32_t n_verts;      /* total vertices over all shapes                       */
    int32_t n_elem_idx;   /* total ints in elem_idx (3 / 2 / 1 per element)        */

    /* per shape [n_shapes]  (src/scene.h:26-50) */
    const int32_t* shape_kind;      /* YRT_TRIANGLES / YRT_LINES / YRT_POINTS        */
    const int32_t* shape_elem_off;  /* first int of this shape in elem_idx           */
    const int32_t* shape_elem_cnt;  /* number of ELEMENTS                            */
    const int32_t* shape_vert_off;  /* first vertex of this shape in pos/norm/…      */
    const int32_t* shape_vert_cnt;
    const int32_t* shape_has_uv;    /* 0: shape::texcoord empty -> uv := (0,0)       */
    const int32_t* shape_has_radius;/* 0: shape::radius empty (triangles)            */

    const int32_t* elem_idx;        /* [n_elem_idx]                                  */
    const float* pos;               /* [3*n_verts]  shape::pos                       */
    const float* norm;              /* [3*n_verts]  shape::norm (tangent for lines)  */
    const float* uv;                /* [2*n_verts]  shape::texcoord, 0 where absent  */
    const float* radius;            /* [n_verts]    shape::radius,   0 where absent  */

    /* per instance [n_instances]  (src/scene.h:99-111), in scn->instances order */
    const float* inst_frame;        /* 12 floats each: x, y, z, o (src/vmath.h:145)  */
    const int32_t* inst_shape;      /* index into shapes                             */
    const int32_t* inst_mat;        /* index into materials                          */

    /* per material [n_materials]  (src/scene.h:62-87) */
    const float* mat_ke;            /* 3 each */
    const float* mat_kd;
    const float* mat_ks;
    const float* mat_kr;
    const float* mat_rs;            /* 1 each; ns is derived on the host like src/raytrace.cpp:144 */
    const int32_t* mat_kd_tex;      /* texture index or -1 */
    const int32_t* mat_ks_tex;

    /* per texture [n_textures]: texture::ldr (src/scene.h:54-58), RGBA8 row-major  */
    const int32_t* tex_w;
    const int32_t* tex_h;
    const int64_t* tex_off;         /* byte offset of texel (0,0) in tex_rgba8       */
    const uint8_t* tex_rgba8;
    int64_t tex_bytes;
} yrt_scene_desc;

/* camera (src/scene.h:115-123); only cameras.front() is used (src/raytrace.cpp:215) */
typedef struct yrt_camera {
    float frame[12];   /* x, y, z, o */
    float fovy, aspect, aperture, focus;
} yrt_camera;

/* per-frame counters and timings filled by the render entry points (optional) */
typedef struct yrt_stats {
    int64_t primary_rays;     /* intersect_first calls the reference would make from raytrace() */
    int64_t reflection_rays;  /* intersect_first calls from recursive shade()                   */
    int64_t shadow_rays;      /* intersect_any calls                                            */
    int64_t launches;         /* kernels of this library launched for the frame                 */
    float ms_total;           /* CUDA-event time of the whole call on the device(s), max        */
    float ms_trace_closest;   /* sum over launches, device 0                                    */
    float ms_trace_any;
    float ms_shade;
    float ms_other;
    float ms_gather;          /* multi-GPU framebuffer gather                                   */
    int32_t max_depth;        /* deepest reflection recursion reached                           */
    int32_t n_gpus;
    int32_t n_closest, n_any, n_shade, n_other;   /* launches per category (device 0)          */
    int32_t frames;           /* frames the totals cover (1, or all frames between yrt_stats_begin/end) */
    int32_t reserved;
    int64_t truncated_paths;  /* mirror bounces dropped at the recursion cap (YRT_MAX_DEPTH, default 64); the
                                 reference recurses without a bound (src/raytrace.cpp:190-204) — 0 means the
                                 frame is what unbounded recursion gives                                      */
} yrt_stats;

typedef struct yrt_scene yrt_scene;   /* opaque: device-resident scene + LBVH on every initialised GPU */

/* ---- library / device management ------------------------------------------------ */
int yrt_abi_version(void);
const char* yrt_last_error(void);
/* number of CUDA devices visible (0 if none / no driver); never fails */
int yrt_device_count(void);
/* use devices 0..n_gpus-1 for the in-process path (yrt_render); n_gpus<=0 means 1.
 * May be called again to change the set while no scene is alive. */
int yrt_init(int n_gpus);
/* use exactly this device (one process per GPU: torchrun ranks pass LOCAL_RANK) */
int yrt_init_device(int device);

/* ---- scene ----------------------------------------------------------------------- */
/* replaces build_bvh(scn,false) (src/raytrace.cpp:278): validates, uploads to every
 * initialised GPU and builds the two-level LBVH there (Morton + radix sort + Karras). */
int yrt_scene_create(const yrt_scene_desc* desc, yrt_scene** out);
void yrt_scene_destroy(yrt_scene* scn);
/* build facts: out[0]=blas nodes, [1]=tlas nodes, [2]=blas max depth, [3]=tlas max depth,
 * [4]=lights, [5]=prims, [6]=build microseconds (device 0), [7]=reflective materials */
int yrt_scene_info(const yrt_scene* scn, int64_t out[8]);

/* Allocates the render workspace of a (width x height, samples per axis) frame now, so that the first yrt_render of that
 * size does not pay for it (1.7 GB for 1920x1080, 16 spp).  Optional. */
int yrt_scene_prepare(yrt_scene* scn, int width, int height, int samples);

/* Number of instances whose frame is not rigid (x, y, z orthonormal within 1e-4).  The reference hands every instance
 * the ray through transform_ray_inverse (src/vmath.h:275-278: dot products with the frame axes, direction re-normalised),
 * which is the inverse only of a rigid frame, and then compares the LOCAL hit distances of different instances with each
 * other and with the world-space boxes (src/scene.cpp:468-473): for scaled or sheared frames (its OBJ `i` lines and glTF node
 * transforms may carry them; none of its own scenes does) what it returns depends on which instances its own BVH lets a
 * ray test, and in which order.  A scene with such frames is therefore traced through a copy of the reference's own
 * instance tree — same nodes, same boxes, same slab test, same stack order — and matches the reference like any other
 * scene; it only runs slower (no LBVH over the instances, no apex grids).  This count tells a caller which case it is in.
 * Pure host function, needs no GPU. */
int yrt_desc_nonrigid_instances(const yrt_scene_desc* desc);
/* process-wide options; unknown names return YRT_ERR_INVALID.
 *   "allow_nonrigid"  accepted and ignored (ABI 2 refused scenes with non-rigid instance frames unless this was set).
 *   "pin_host_frames" (0/1, default 0): yrt_render / yrt_render_ldr page-lock the caller's output buffer the first time they
 *                     see it and keep it registered while the same buffer keeps coming back (device->host copies into
 *                     pageable memory are staged by the driver).  The registration outlives the call: only for callers
 *                     that keep the buffer alive, like the CLI. */
int yrt_set_option(const char* name, int value);

/* Optional replacement of save_image(filename, image4b) (src/image.cpp:41-44, stb_image_write's single-threaded PNG
 * encoder: 0.63 s at 1920x1080): writes the same RGBA8 pixels (row-major, width*4 bytes per row, like
 * image4b) as a PNG encoded on `threads` host threads (0 = all cores; level = zlib level 1..9, 0 = default 1).  PNG is
 * lossless, so the decoded image is identical to the reference's file; the file bytes are not.  Host only, needs no GPU. */
int yrt_write_png(const char* path, const uint8_t* rgba8, int width, int height, int threads, int level);

/* image width the reference derives from the camera: (int)std::round(aspect*resolution)
 * (src/raytrace.cpp:216) */
int yrt_image_width(const yrt_camera* cam, int resolution);

/* ---- render ---------------------------------------------------------------------- */
/* replaces raytrace(scn, amb, resolution, samples) (src/raytrace.cpp:213-254).
 * `samples` is the per-axis count N (N*N samples per pixel, src/raytrace.cpp:232-234).
 * rgba_out: HOST buffer, width*height*4 floats, row-major pixels[j*width+i] like image4f
 * (src/image.h:15).  Uses every GPU given to yrt_init: interleaved row tiles, and every GPU copies its
 * own rows into rgba_out over its own PCIe link (no exchange between GPUs); see "pin_host_frames" for
 * pageable output buffers. */
int yrt_render(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height,
               int samples, float* rgba_out, yrt_stats* stats);

/* Next row: the same frame followed by tonemap(hdr, 0, false) (src/image.cpp:55-78) ON THE DEVICE, i.e. what
 * save_hdr_or_ldr hands to the PNG writer: rgba8_out = width*height*4 bytes (HOST).  Only a quarter of the bytes cross to the
 * host.  rgba_out (optional, HOST floats) additionally receives the float image.  The device powf differs from glibc's by a few
 * ulp: at most one 8-bit level on ~0.3 % of pixels against the host tonemap. */
int yrt_render_ldr(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples,
                   uint8_t* rgba8_out, float* rgba_out, yrt_stats* stats);

/* Same frame, device-resident, for one-process-per-GPU drivers: renders only the rows
 * owned by `rank` of `world` (row tile t of `tile_rows` rows belongs to rank t % world)
 * into d_rgba (DEVICE pointer on the current yrt device) holding the rank's rows packed in
 * increasing row order: yrt_rows_owned(height,tile_rows,rank,world)*width*4 floats.
 * stream: cudaStream_t as void*; NULL = the legacy default stream (what PyTorch's default stream is), so the
 * frame is ordered with the caller's other work on that stream.  Asynchronous w.r.t. the host unless
 * stats != NULL (and no deferred statistics are open). */
int yrt_render_rows(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height,
                    int samples, int tile_rows, int rank, int world, void* d_rgba, void* stream,
                    yrt_stats* stats);
int yrt_rows_owned(int height, int tile_rows, int rank, int world);
/* Same rows, delivered to their final positions of a row-major HOST frame of the whole image (width*height*4 floats) —
 * for one-process-per-GPU drivers whose ranks share one host frame (POSIX shared memory, page-locked by every rank):
 * each rank moves only its own rows, over its own PCIe link, with one pitched copy; the ranks exchange nothing.
 * Asynchronous on `stream` unless stats != NULL.  The caller orders frames: rank r may overwrite its rows of frame k
 * as soon as it is called for frame k+1, so the consumer must be done with frame k by then (a barrier, or two frames). */
int yrt_render_rows_to_host(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height,
                            int samples, int tile_rows, int rank, int world, float* h_frame, void* stream,
                            yrt_stats* stats);
/* Fused gather for one-process-per-GPU drivers: rank 0 allocates the full frame (yrt_frame_alloc) and exports a
 * 64-byte CUDA IPC handle; the other ranks map it (yrt_frame_import: peer mapping over NVLink) and every rank's
 * resolve kernel stores its rows at their final position in that ONE buffer (yrt_render_rows_into_frame) — no
 * packed rows, no gather copy, no unpack.  The allocation holds TWO frames back to back (frame k goes to d_full +
 * (k % 2) * width*height*16 bytes) and an arrival counter behind them; every frame ends with ONE barrier:
 * yrt_frame_barrier(base, width, height, world, k + 1, stream) enqueues a one-thread kernel that fences this rank's
 * stores, adds its arrival to the counter in rank 0's memory (a peer atomic over NVLink) and waits until all `world`
 * ranks of frame k have arrived — no collective library call, no host synchronisation.  That orders everything: rank 0
 * may read frame k once its own barrier kernel has completed (stream order); a rank writes into frame k's buffer
 * again only in frame k + 2, i.e. after barrier k + 1, which rank 0 joins after the reads of frame k it enqueued on
 * the same stream (a consumer that reads on another stream or on the host finishes before it calls frame k + 1).
 * generation counts frames from 1 and only grows (distributed.SharedFrame does all of this). */
int yrt_render_rows_into_frame(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height,
                               int samples, int tile_rows, int rank, int world, void* d_full, void* stream,
                               yrt_stats* stats);
int yrt_frame_alloc(int width, int height, void** d_full);
int yrt_frame_free(void* d_full);
int yrt_frame_export(void* d_full, unsigned char handle[64]);
int yrt_frame_import(const unsigned char handle[64], void** d_full);
int yrt_frame_release(void* d_full);
int yrt_frame_barrier(void* d_full_base, int width, int height, int world, int64_t generation, void* stream);
/* Barrier between the processes of one node through a counter in memory they share (8 bytes, zero-initialised, e.g. the tail of
 * the shared host frame): adds one arrival and returns when world * generation arrivals have been counted; generation =
 * 1, 2, 3, ... (the counter only grows).  Microseconds instead of a collective launch + stream wait per frame.  Host only. */
int yrt_host_barrier(void* counter, int world, int64_t generation);

/* deferred statistics for yrt_render_rows*: frames rendered between begin and end record per-launch CUDA events
 * and ray counters without waiting for the device; yrt_stats_end waits for it and returns the totals over those
 * frames (stats->frames says how many).  (Scenes with mirrors: the host waits, one wave behind the device, for the
 * size of each reflection wave — see yrt_render.cu; the device is never idle because of it.) */
int yrt_stats_begin(yrt_scene* scn);
int yrt_stats_end(yrt_scene* scn, yrt_stats* totals);
/* scatter a rank's packed rows into the full row-major framebuffer (both DEVICE pointers) */
int yrt_unpack_rows(const void* d_packed, void* d_full, int width, int height, int tile_rows,
                    int rank, int world, void* stream);

/* parity hook for the closest-hit metric: for every primary ray, in the order
 * ((j*width+i)*samples+jj)*samples+ii, ids_out[3*r+0..2] = (instance index in
 * scn->instances, shape index in scn->shapes, element index ei) or (-1,-1,-1) on a miss;
 * dist_out[r] = hit distance (intersection3f::dist, src/scene.h:231) or 0.  HOST buffers;
 * dist_out / uv_out (2 floats per ray: ew.y, ew.z for triangles, ew.y for lines) may be NULL. */
int yrt_trace_primary(yrt_scene* scn, const yrt_camera* cam, int width, int height, int samples,
                      int32_t* ids_out, float* dist_out, float* uv_out);

/* generic ray queries (intersect_first / intersect_any, src/scene.cpp:483-494) on HOST arrays
 * of n rays: rays = 8 floats each (o.xyz, d.xyz, tmin, tmax).  ids_out as above (closest) or
 * occluded_out[n] bytes (any). */
int yrt_intersect_first(yrt_scene* scn, const float* rays, int64_t n, int32_t* ids_out,
                        float* dist_out, float* uv_out);
int yrt_intersect_any(yrt_scene* scn, const float* rays, int64_t n, uint8_t* occluded_out);

/* Mirror bounces the LAST frame of yrt_render / yrt_render_ldr dropped at the recursion cap, summed over the GPUs — the same
 * number as yrt_stats.truncated_paths, for callers that render without per-call statistics (a frame without statistics runs as
 * two overlapping pipelines and reads nothing back).  Waits for the devices; one 40-byte read per GPU.  0 = the frame is what
 * the reference's unbounded recursion (src/raytrace.cpp:190-204) gives.  The drop-in CLI warns whenever it is not 0. */
int yrt_frame_truncated_paths(yrt_scene* scn, int64_t* out);

/* Per-ray work counters of the traversal kernels, for the roofline record ("the builder must also report its
 * own per-ray counters"; the reference's sit at src/scene.cpp:371,229,468).  Only a library built with -DYRT_COUNTERS=1
 * counts (tools/build_variants.sh; never the timed build): the regular build returns YRT_ERR_UNSUPPORTED.  out = 3 kernel
 * classes (camera rays, mirror rays, shadow rays) x 8 words: rays, node visits, box tests, box tests in the instance tree,
 * element tests, instance entries, 0, 0 — totals since the last call (reading resets them). */
int yrt_counters_read(yrt_scene* scn, uint64_t out[24]);

/* test hook: copies the scene's node records (arity 2: 4 float4 per node, arity 4: 8 float4 per node; BLAS nodes of all
 * shapes, then TLAS nodes) of device 0 into out (HOST, capacity max_float4 float4s); returns the number of float4s the
 * array holds (also when out is NULL), or a negative yrt_status.  The LBVH build is deterministic — its lock-free bottom-up
 * passes included — so these arrays must equal, bit for bit, those of a serial execution of the same per-item functions
 * (tests/host_emu): that comparison is the race detector of the build (tests/test_gpu_parity.py). */
int64_t yrt_debug_read_nodes(yrt_scene* scn, int arity, float* out, int64_t max_float4);

/* test hook: sorts n (key, value) pairs held in HOST arrays with the device radix sort of the LBVH build */
int yrt_debug_sort_pairs(unsigned long long* h_keys, int* h_vals, int n);

/* next-row: tonemap (src/image.cpp:55-78, exposure 0, no filmic, sRGB 1/2.2,
 * truncating) of a HOST float image into HOST RGBA8, computed on the device. */
int yrt_tonemap(const float* rgba_in, int width, int height, uint8_t* rgba8_out);

#ifdef __cplusplus
}
#endif
#endif /* YRT_B200_H_ */
