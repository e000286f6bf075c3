/*
 * yrt_oracle.h — TEST INFRASTRUCTURE.  CPU restatement (plain C) of the reference's render path on the
 * flattened scene of include/yrt_b200.h.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product (libyrt_b200.so) never does.
 *
 * Pinned: tests/test_oracle.py checks it against outputs of the UNMODIFIED reference compiled here
 * (oracle/_ref, built from /root/reference/src by the Makefile) — committed under tests/golden/ with the
 * generating script tools/make_golden.py — bit-exact on the float image and on every hit id.
 */
#ifndef YRT_ORACLE_H_
#define YRT_ORACLE_H_
#include <stdint.h>

#include "../include/yrt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_scene oracle_scene;

/* build_bvh(scn,false), src/scene.cpp:554-565 (+ :525-549, :572-658) */
int oracle_scene_create(const yrt_scene_desc* desc, oracle_scene** out);
void oracle_scene_destroy(oracle_scene* scn);
/* out[0]=scene-bvh nodes, out[1]=total shape-bvh nodes, out[2]=lights */
void oracle_scene_info(const oracle_scene* scn, int64_t out[4]);

/* raytrace(), src/raytrace.cpp:213-254.  counts (optional): intersect_first calls from raytrace(),
 * intersect_first calls from recursive shade(), intersect_any calls, max recursion depth.
 * n_threads > 1 splits image rows over OpenMP threads (pixels are independent; the reference itself is
 * single-threaded); max_depth <= 0 means unbounded like the reference. */
int oracle_render(const oracle_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples,
                  int max_depth, int n_threads, float* rgba_out, int64_t counts[4]);
/* rows [row0, row1) only (bounded CPU-baseline samples); rgba_out still indexes the full image */
int oracle_render_rows(const oracle_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples,
                       int max_depth, int n_threads, int row0, int row1, float* rgba_out, int64_t counts[4]);

/* intersect_first per primary ray, same order and id definitions as yrt_trace_primary */
int oracle_trace_primary(const oracle_scene* scn, const yrt_camera* cam, int width, int height, int samples, int brute_force,
                         int32_t* ids_out, float* dist_out, float* uv_out);
int oracle_intersect_first(const oracle_scene* scn, const float* rays, int64_t n, int32_t* ids_out, float* dist_out, float* uv_out);
int oracle_intersect_any(const oracle_scene* scn, const float* rays, int64_t n, uint8_t* occluded_out);

/* tonemap(hdr, 0, false), src/image.cpp:55-78 */
void oracle_tonemap(const float* rgba_in, int width, int height, uint8_t* rgba8_out);
int oracle_image_width(const yrt_camera* cam, int resolution);

#ifdef __cplusplus
}
#endif
#endif
