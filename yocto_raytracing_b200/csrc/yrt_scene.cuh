// yrt_scene.cuh — device-resident scene layout (what lives in HBM) and node/leaf references.
//
// All records are multiples of 16 bytes and are fetched with 128-bit loads.
//   BVH node (TLAS and BLAS alike, boxes as centre / half-extent inflated by 1 + 16u, two children interleaved per
//   48-byte "pair" so that the slab test runs on packed FFMA2 — see "traversal node" below):
//     4-wide: 128 B = pair(children 0,1) | pair(children 2,3) | 4 child refs | child count
//     binary:  64 B = pair(children 0,1) | 2 child refs            (both arrays are built; see YRT_WIDE_CLOSEST / YRT_WIDE_ANY)
//   prim record (48 B, in BLAS leaf order, all shapes concatenated):
//     triangle: q0 = v0.xyz | ei     q1 = e1.xyz | -      q2 = e2.xyz | -     (e1 = v1 - v0, e2 = v2 - v0: scene.cpp:236-237, subtracted once at build time)
//     line:     q0 = v0.xyz | ei     q1 = v1.xyz | r0     q2 = r1, -, -, -
//     point:    q0 = p.xyz  | ei     q1 = r, -, -, -
//   prim attribute record (YRT_ATTR_STRIDE quads, same order; pre-gathered shape::norm / shape::texcoord, and a triangle's v1, v2 for eval_pos):
//     q0 = n0.xyz | uv0.x   q1 = n1.xyz | uv0.y   q2 = n2.xyz | uv1.x   q3 = uv1.y, uv2.x, uv2.y, -   q4 = v1.xyz   q5 = v2.xyz
//   instance record (64 B, in TLAS leaf order):
//     q0 = frame.x | blas root ref   q1 = frame.y | instance index (scn->instances order)
//     q2 = frame.z | material index  q3 = frame.o | shape index | kind << 28
//   material record (64 B): kd.xyz|ns  ks.xyz|kd_tex  kr.xyz|ks_tex  ke.xyz|-
//   light record (80 B): frame x,y,z,o (4 x float4, .w unused) + q4 = pos0.xyz | -,  then ke in q0..q2 .w
#pragma once
#include "yrt_math.cuh"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
#include <vector_types.h>
#endif

namespace yrt {

// ---- references held in node child slots and on the traversal stack -----------------------
//   ref >= 0         : internal node index
//   ref <  0         : leaf; ~ref = (first << 3) | (count - 1), 1 <= count <= 8
//   YRT_REF_SENTINEL : traversal stack marker "leave the current instance"
#define YRT_REF_SENTINEL ((int)0x80000000)
#define YRT_REF_DONE ((int)0x80000001)      /* traversal finished (never a valid leaf: first <= YRT_MAX_LEAF_FIRST) */
#define YRT_LEAF_MAX_COUNT 8
#define YRT_MAX_LEAF_FIRST ((1 << 28) - 2)
#define YRT_ATTR_STRIDE 6
#define YRT_STACK_CAP 128   /* traversal stack entries (TLAS + BLAS levels simultaneously live); checked against the built depth */

YRT_HD int make_leaf_ref(int first, int count) { return ~((first << 3) | (count - 1)); }
YRT_HD int leaf_first(int ref) { return (~ref) >> 3; }
YRT_HD int leaf_count(int ref) { return ((~ref) & 7) + 1; }

YRT_HD float int_as_float(int i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    union { int i; float f; } u; u.i = i; return u.f;
#endif
}
YRT_HD int float_as_int(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    union { int i; float f; } u; u.f = f; return u.i;
#endif
}

YRT_HD float4 ld4(const float4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

YRT_HD float4 mk4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- traversal node ---------------------------------------------------------------------------
// A "pair" is the 48-byte record of two sibling child boxes:
//   p0 = (c0.x, c0.y, c1.x, c1.y)   p1 = (h0.x, h0.y, h1.x, h1.y)   p2 = (c0.z, c1.z, h0.z, h1.z)
// The two boxes are interleaved so that every FMA of the slab test has a partner with which it shares one packed
// FFMA2 (sm_100: two fp32 FMAs per issue slot on an aligned register pair): x and y of one child pair up against the
// ray's (invd.x, invd.y) / (noi.x, noi.y) / (ainv.x, ainv.y) pairs, and the z of child 0 pairs with the z of child 1
// against a broadcast scalar.  20 FFMA per pair become 10 FFMA2; the values are the same IEEE FMAs, so the accept
// decisions are bit-identical to slab_test_ch on the same boxes (the host build below uses exactly that).
//
// Two node arrays are built from the same binary LBVH and share its node indices (record i of either array belongs to
// binary node i, so child references and roots are the same numbers in both):
//   4-wide: 8 float4, 128-byte aligned = one L1 line: pair(0,1) | pair(2,3) | (ref0, ref1, ref2, ref3) | (children, -, -, -).
//           A visit tests up to four (grand)children of a binary node at once: half as many dependent fetches and loop
//           trips; unused child slots hold a box no ray can enter (h = -FLT_MAX).  Collapsed at emit time (yrt_lbvh.cuh).
//   binary: 4 float4: pair(0,1) | (ref0, ref1, -, -).
// Which kernel walks which array is a build-time choice per ray kind, measured on B200 (profiles/r2_experiments.md):
// any-hit (shadow) rays are faster on the 4-wide nodes (no child ordering: the visit is lean, the halved trip count
// pays), closest-hit rays on the binary ones (ordering four children costs more than the saved trips).
#ifndef YRT_WIDE_CLOSEST
#define YRT_WIDE_CLOSEST 2
#endif
#ifndef YRT_WIDE_ANY
#define YRT_WIDE_ANY 4
#endif
#define YRT_NODE_STRIDE(W) ((W) == 4 ? 8 : 4)   /* float4 per node record */

struct nodebox { float cx, cy, cz, hx, hy, hz; };
YRT_HD nodebox pair_child(const float4& p0, const float4& p1, const float4& p2, int k) {
    nodebox b;
    if (k == 0) { b.cx = p0.x; b.cy = p0.y; b.cz = p2.x; b.hx = p1.x; b.hy = p1.y; b.hz = p2.z; }
    else        { b.cx = p0.z; b.cy = p0.w; b.cz = p2.y; b.hx = p1.z; b.hy = p1.w; b.hz = p2.w; }
    return b;
}
// a box no ray enters: t0 = +huge, t1 = -huge on every axis (fused test), lo > hi (reference formula)
YRT_HD nodebox nodebox_empty() { nodebox b; b.cx = b.cy = b.cz = 0.f; b.hx = b.hy = b.hz = -FLT_MAX; return b; }

YRT_HD void pair_pack(float4* p, const nodebox& b0, const nodebox& b1) {
    p[0] = mk4(b0.cx, b0.cy, b1.cx, b1.cy);
    p[1] = mk4(b0.hx, b0.hy, b1.hx, b1.hy);
    p[2] = mk4(b0.cz, b1.cz, b0.hz, b1.hz);
}

// writes one node record of arity W: nb child boxes (2 <= nb <= W) and their references; the remaining slots are empty
template <int W>
YRT_HD void node_pack(float4* n, const nodebox* box, const int* ref, int nb) {
    nodebox b[W];
    int r[W];
    for (int k = 0; k < W; k++) {
        b[k] = k < nb ? box[k] : nodebox_empty();
        r[k] = k < nb ? ref[k] : YRT_REF_DONE;   // never read: the slot's box cannot be entered
    }
    pair_pack(n, b[0], b[1]);
    if (W == 4) {
        pair_pack(n + 3, b[2], b[3]);
        n[6] = mk4(int_as_float(r[0]), int_as_float(r[1]), int_as_float(r[W - 2]), int_as_float(r[W - 1]));
        n[7] = mk4(int_as_float(nb), 0.f, 0.f, 0.f);
    } else {
        n[3] = mk4(int_as_float(r[0]), int_as_float(r[1]), 0.f, 0.f);
    }
}
// child k of a node record held in host-visible memory (build passes, tools and tests; the traversal reads the quads itself)
template <int W>
YRT_HD void node_child(const float4* n, int k, nodebox& b, int& ref) {
    const float4* p = n + 3 * (k >> 1);
    b = pair_child(p[0], p[1], p[2], k & 1);
    const float4 q = n[W == 4 ? 6 : 3];
    ref = float_as_int(k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w)));
}

#if defined(__CUDA_ARCH__)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#endif

// both children of one pair against one ray; e0/e1 = entry distances (ordering only)
YRT_HD void slab_test_pair(const slabray& r, float rtmin, float rtmax, const float4& q0, const float4& q1, const float4& q2,
                           bool& h0, bool& h1, float& e0, float& e1) {
#if defined(__CUDA_ARCH__) && !defined(YRT_NO_FFMA2)
    f32x2 ixy = pk2(r.invd.x, r.invd.y), nxy = pk2(r.noi.x, r.noi.y), axy = pk2(r.ainv.x, r.ainv.y), naxy = pk2(-r.ainv.x, -r.ainv.y);
    f32x2 izz = pk2(r.invd.z, r.invd.z), nzz = pk2(r.noi.z, r.noi.z), azz = pk2(r.ainv.z, r.ainv.z), nazz = pk2(-r.ainv.z, -r.ainv.z);
    f32x2 tc0 = fma2(pk2(q0.x, q0.y), ixy, nxy), tc1 = fma2(pk2(q0.z, q0.w), ixy, nxy), tcz = fma2(pk2(q2.x, q2.y), izz, nzz);
    f32x2 hh0 = pk2(q1.x, q1.y), hh1 = pk2(q1.z, q1.w), hhz = pk2(q2.z, q2.w);
    f32x2 lo0 = fma2(hh0, naxy, tc0), hi0 = fma2(hh0, axy, tc0);
    f32x2 lo1 = fma2(hh1, naxy, tc1), hi1 = fma2(hh1, axy, tc1);
    f32x2 loz = fma2(hhz, nazz, tcz), hiz = fma2(hhz, azz, tcz);
    float t0x0, t0y0, t1x0, t1y0, t0x1, t0y1, t1x1, t1y1, t0z0, t0z1, t1z0, t1z1;
    upk2(lo0, t0x0, t0y0); upk2(hi0, t1x0, t1y0); upk2(lo1, t0x1, t0y1); upk2(hi1, t1x1, t1y1);
    upk2(loz, t0z0, t0z1); upk2(hiz, t1z0, t1z1);
    float tmin0 = fmaxf(t0z0, fmaxf(t0y0, fmaxf(t0x0, rtmin))), tmax0 = fminf(t1z0, fminf(t1y0, fminf(t1x0, rtmax)));
    float tmin1 = fmaxf(t0z1, fmaxf(t0y1, fmaxf(t0x1, rtmin))), tmax1 = fminf(t1z1, fminf(t1y1, fminf(t1x1, rtmax)));
    float lim0, lim1;
    upk2(fma2(pk2(tmax0, tmax1), pk2(YRT_SLAB_ACCEPT, YRT_SLAB_ACCEPT), pk2(r.pad, r.pad)), lim0, lim1);
    e0 = tmin0; e1 = tmin1;
    h0 = tmin0 <= lim0; h1 = tmin1 <= lim1;
#else
    h0 = slab_test_ch(r, rtmin, rtmax, q0.x, q0.y, q2.x, q1.x, q1.y, q2.z, e0);
    h1 = slab_test_ch(r, rtmin, rtmax, q0.z, q0.w, q2.y, q1.z, q1.w, q2.w, e1);
#endif
}

YRT_HD vec3 xyz(const float4& q) { return mk3(q.x, q.y, q.z); }

// ---- what a kernel sees ------------------------------------------------------------------
struct SceneView {
    const float4* nodes2;       // binary node records (4 float4 each): BLAS nodes of all shapes first, TLAS nodes after them —
                                // one array and one index space, so a node visit needs no level test
    const float4* nodes4;       // the same tree as 4-wide records (8 float4 each), same indices
    const float4* inst_recs;    // 4 per instance, TLAS leaf order
    const float4* inst_box;     // 2 per instance, same order: world box as (centre, inflated half-extent) — what the apex grids are built from (yrt_pgrid.cuh)
    const float4* prim_recs;    // 3 per prim, BLAS leaf order
    const float4* prim_attrs;   // YRT_ATTR_STRIDE per prim: normals + uv (4), triangles: v1, v2 (the trace record holds edges)
    const float4* mat_recs;     // 4 per material
    const float4* light_recs;   // 5 per light
    const uint8_t* tex_rgba8;   // all textures
    const int4* tex_info;       // per texture: w, h, byte offset lo, byte offset hi
    const float* srgb_lut;      // 256 floats: fminf(1, powf(b/255, 2.2f)) computed with the HOST libm
    // exact-distance ties: position of each instance / element in the reference's own (ray-independent)
    // BVH visit sequence — the later-visited candidate wins a tie there (see yrt_host.cu); read on ties only
    const int* inst_rank;       // per instance slot (TLAS leaf order)
    const int* prim_rank;       // per prim slot (BLAS leaf order), rank inside its shape
    int tlas_root;              // ref
    int n_lights;
    int n_active_instances;
};

// RefTlas — only for scenes with scaled / sheared instance frames (n_nodes > 0), which the LBVH cannot serve:
// transform_ray_inverse (vmath.h:275-278) inverts rigid frames only, so for any other frame the geometry the reference
// "sees" is not inside the instance's world box, and the local hit distance it keeps as tray.tmax (scene.cpp:470) is not
// a world distance.  What the reference returns then depends on WHICH instances its own tree lets a ray test and in which
// order.  Such scenes are traced through a copy of that tree (yrt_host.cu: visit_ranks): same nodes, same boxes, same slab
// formula, same stack order (trace_ray_ref, yrt_trace.cuh); the shapes' trees below stay the LBVH's.  Kept out of SceneView
// so that the kernels of every other scene do not change by a byte.
struct RefTlas {
    const float4* nodes;        // 2 per node: {bbox.min | a}, {bbox.max | b}; inner: a, b = children; leaf: a = ~first, b = count
    const int* leaf_inst;       // the reference's leaf_prims: instance ids
    const int* slot_of_inst;    // instance id -> slot in inst_recs (-1: its shape has no elements)
    int n_nodes;
};

// trace record (3 float4) + the vertex positions shading needs, for one element in BLAS leaf order.
//   triangle: (v0 | element, e1 = v1 - v0, e2 = v2 - v0)  — the edges of scene.cpp:236-237, subtracted once here
//   line:     (v0 | element, v1 | r0, r1)        point: (v0 | element, r0)
// ar[4], ar[5] keep v1, v2 of a triangle for eval_hit (scene.h:166-168 interpolates the vertices, not the edges)
YRT_HD void pack_prim(int kind, int e, const vec3& p0, const vec3& p1, const vec3& p2, float r0, float r1, float4* pr, float4* ar) {
    pr[0] = mk4(p0.x, p0.y, p0.z, int_as_float(e));
    if (kind == 0) {
        vec3 e1 = p1 - p0, e2 = p2 - p0;
        pr[1] = mk4(e1.x, e1.y, e1.z, 0.f);
        pr[2] = mk4(e2.x, e2.y, e2.z, 0.f);
    } else if (kind == 1) {
        pr[1] = mk4(p1.x, p1.y, p1.z, r0);
        pr[2] = mk4(r1, 0.f, 0.f, 0.f);
    } else {
        pr[1] = mk4(r0, 0.f, 0.f, 0.f);
        pr[2] = mk4(0.f, 0.f, 0.f, 0.f);
    }
    ar[4] = mk4(p1.x, p1.y, p1.z, 0.f);
    ar[5] = mk4(p2.x, p2.y, p2.z, 0.f);
}

// closest-hit result (16 B on the wire: si, prim, w1, w2; dist kept for the parity hook)
struct HitRec {
    int si;      // instance slot in TLAS leaf order (-1 = miss)
    int prim;    // prim slot in BLAS leaf order
    float w1;    // triangle: ew.y ; line: s (= ew.y) ; point: 0
    float w2;    // triangle: ew.z
    float dist;
};

}  // namespace yrt
