/*
 * yrt_oracle.c — TEST INFRASTRUCTURE (see yrt_oracle.h).  Plain-C restatement of the reference's hot
 * path, function by function, on the flattened scene.  Citations are to /root/reference/src.
 * Compile WITHOUT -march=native / -ffast-math and with -ffp-contract=off: the reference is built for
 * baseline x86-64 (no FMA), and hit ids depend on that (SURVEY.md finding 4).
 * Nothing here is shared with the CUDA path: different BVH (the reference's top-down midpoint split,
 * 32-byte nodes, unordered two-child push), different traversal, recursive shade().
 */
#include "yrt_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- vmath.h ---------------------------------------------------------------------------- */
typedef struct { float x, y; } vec2f;
typedef struct { float x, y, z; } vec3f;
typedef struct { float x, y, z, w; } vec4f;
typedef struct { vec3f x, y, z, o; } frame3f;                 /* vmath.h:145 */
typedef struct { vec3f o, d; float tmin, tmax; } ray3f;       /* vmath.h:266-271 */
typedef struct { vec3f min, max; } bbox3f;                    /* vmath.h:280 */
static const float ray_eps = 1e-4f;                           /* vmath.h:264 */

static inline vec3f v3(float x, float y, float z) { vec3f r = {x, y, z}; return r; }
static inline vec3f add3(vec3f a, vec3f b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }          /* :69 */
static inline vec3f sub3(vec3f a, vec3f b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }          /* :73 */
static inline vec3f mulf3(vec3f a, float b) { return v3(a.x * b, a.y * b, a.z * b); }               /* :77 */
static inline vec3f divf3(vec3f a, float b) { return v3(a.x / b, a.y / b, a.z / b); }               /* :81 */
static inline vec3f mul3(vec3f a, vec3f b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }          /* :85 */
static inline vec2f add2(vec2f a, vec2f b) { vec2f r = {a.x + b.x, a.y + b.y}; return r; }          /* :61 */
static inline vec2f mulf2(vec2f a, float b) { vec2f r = {a.x * b, a.y * b}; return r; }             /* :65 */
static inline float dot3(vec3f a, vec3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            /* :112 */
static inline float length3(vec3f a) { return sqrtf(dot3(a, a)); }                                  /* :116 */
static inline vec3f normalize3(vec3f a) {                                                           /* :118-122 */
    float l = length3(a);
    if (l == 0) return a;
    return mulf3(a, 1 / l);
}
static inline vec3f cross3(vec3f a, vec3f b) {                                                      /* :124-127 */
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float minf_(float x, float y) { return (x < y) ? x : y; }                             /* :215 */
static inline float maxf_(float x, float y) { return (x > y) ? x : y; }                             /* :216 */
static inline float clampf_(float x, float a, float b) { return minf_(maxf_(x, a), b); }            /* :217 */

static inline vec3f transform_point(const frame3f* a, vec3f b) {                                    /* :152 */
    return add3(add3(add3(mulf3(a->x, b.x), mulf3(a->y, b.y)), mulf3(a->z, b.z)), a->o);
}
static inline vec3f transform_vector(const frame3f* a, vec3f b) {                                   /* :161 */
    return add3(add3(mulf3(a->x, b.x), mulf3(a->y, b.y)), mulf3(a->z, b.z));
}
static inline vec3f transform_direction(const frame3f* a, vec3f b) { return normalize3(transform_vector(a, b)); }   /* :169 */
static inline vec3f transform_point_inverse(const frame3f* a, vec3f b) {                            /* :156-159 */
    vec3f bo = sub3(b, a->o);
    return v3(dot3(a->x, bo), dot3(a->y, bo), dot3(a->z, bo));
}
static inline vec3f transform_direction_inverse(const frame3f* a, vec3f b) {                        /* :165-175 */
    return normalize3(v3(dot3(a->x, b), dot3(a->y, b), dot3(a->z, b)));
}
static inline vec3f eval_ray(const ray3f* ray, float t) { return add3(ray->o, mulf3(ray->d, t)); }  /* :273 */
static inline ray3f transform_ray_inverse(const frame3f* f, const ray3f* ray) {                     /* :275-278 */
    ray3f r;
    r.o = transform_point_inverse(f, ray->o);
    r.d = transform_direction_inverse(f, ray->d);
    r.tmin = ray->tmin;
    r.tmax = ray->tmax;
    return r;
}
static const bbox3f invalid_bbox3f = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};   /* :284 */
static inline bbox3f expand_bbox_p(bbox3f a, vec3f b) {                                             /* :287-290 */
    bbox3f r = {{minf_(a.min.x, b.x), minf_(a.min.y, b.y), minf_(a.min.z, b.z)},
                {maxf_(a.max.x, b.x), maxf_(a.max.y, b.y), maxf_(a.max.z, b.z)}};
    return r;
}
static inline bbox3f expand_bbox_b(bbox3f a, bbox3f b) {                                            /* :292-296 */
    bbox3f r = {{minf_(a.min.x, b.min.x), minf_(a.min.y, b.min.y), minf_(a.min.z, b.min.z)},
                {maxf_(a.max.x, b.max.x), maxf_(a.max.y, b.max.y), maxf_(a.max.z, b.max.z)}};
    return r;
}
static bbox3f bbox_to_world(const frame3f* a, bbox3f b) {                                           /* :312-326 */
    vec3f corners[8] = {
        {b.min.x, b.min.y, b.min.z}, {b.min.x, b.min.y, b.max.z}, {b.min.x, b.max.y, b.min.z}, {b.min.x, b.max.y, b.max.z},
        {b.max.x, b.min.y, b.min.z}, {b.max.x, b.min.y, b.max.z}, {b.max.x, b.max.y, b.min.z}, {b.max.x, b.max.y, b.max.z},
    };
    bbox3f bbox = invalid_bbox3f;
    for (int i = 0; i < 8; i++) bbox = expand_bbox_p(bbox, transform_point(a, corners[i]));
    return bbox;
}

/* ---- scene.h ---------------------------------------------------------------------------- */
typedef struct {                                              /* scene.h:9-15 */
    bbox3f bbox;
    uint32_t start;
    uint16_t count;
    uint8_t isleaf;
    uint8_t axis;
} bvh_node;
typedef struct { bvh_node* nodes; int n_nodes, cap_nodes; int* leaf_prims; int n_prims; } bvh_tree;   /* scene.h:19-22 */

typedef struct {                                              /* scene.h:26-50, flattened */
    int kind, n_elems, n_verts, has_uv;
    const int32_t* elems;
    const float *pos, *norm, *uv, *radius;
    bvh_tree bvh;
} o_shape;
typedef struct { int w, h; const uint8_t* px; } o_texture;   /* texture::ldr, scene.h:54-58 */
typedef struct { vec3f ke, kd, ks, kr; float rs; int kd_txt, ks_txt; } o_material;   /* scene.h:62-87 */
typedef struct { frame3f frame; int mat, shp; } o_instance;  /* scene.h:99-111 */

struct oracle_scene {
    int n_shapes, n_instances, n_materials, n_textures;
    o_shape* shapes;
    o_instance* instances;
    o_material* materials;
    o_texture* textures;
    bvh_tree bvh;
    /* private copies of the description's arrays */
    int32_t* elem_idx; float *pos, *norm, *uv, *radius; uint8_t* texels;
};

typedef struct { int ist; int ei; vec4f ew; float dist; } intersection3f;   /* scene.h:227-234; ist = instance index, -1 */

static inline vec3f pos_of(const o_shape* s, int v) { return v3(s->pos[3 * v], s->pos[3 * v + 1], s->pos[3 * v + 2]); }
static inline vec3f norm_of(const o_shape* s, int v) { return v3(s->norm[3 * v], s->norm[3 * v + 1], s->norm[3 * v + 2]); }
static inline vec2f uv_of(const o_shape* s, int v) {
    vec2f r = {0.f, 0.f};   /* reference indexes an empty vector here (UB, scene.h:193-205); defined as (0,0) */
    if (s->has_uv) { r.x = s->uv[2 * v]; r.y = s->uv[2 * v + 1]; }
    return r;
}

/* scene.h:159-172 (points > lines > triangles; every shape holds one kind) */
static vec3f eval_pos_shape(const o_shape* shp, int ei, vec4f ew) {
    if (shp->kind == YRT_POINTS) return pos_of(shp, shp->elems[ei]);
    if (shp->kind == YRT_LINES)
        return add3(mulf3(pos_of(shp, shp->elems[2 * ei]), ew.x), mulf3(pos_of(shp, shp->elems[2 * ei + 1]), ew.y));
    return add3(add3(mulf3(pos_of(shp, shp->elems[3 * ei]), ew.x), mulf3(pos_of(shp, shp->elems[3 * ei + 1]), ew.y)),
                mulf3(pos_of(shp, shp->elems[3 * ei + 2]), ew.z));
}
/* scene.h:176-189 */
static vec3f eval_norm_shape(const o_shape* shp, int ei, vec4f ew) {
    if (shp->kind == YRT_POINTS) return norm_of(shp, shp->elems[ei]);
    if (shp->kind == YRT_LINES)
        return normalize3(add3(mulf3(norm_of(shp, shp->elems[2 * ei]), ew.x), mulf3(norm_of(shp, shp->elems[2 * ei + 1]), ew.y)));
    return normalize3(add3(add3(mulf3(norm_of(shp, shp->elems[3 * ei]), ew.x), mulf3(norm_of(shp, shp->elems[3 * ei + 1]), ew.y)),
                           mulf3(norm_of(shp, shp->elems[3 * ei + 2]), ew.z)));
}
/* scene.h:193-206 */
static vec2f eval_texcoord_shape(const o_shape* shp, int ei, vec4f ew) {
    if (shp->kind == YRT_POINTS) return uv_of(shp, shp->elems[ei]);
    if (shp->kind == YRT_LINES) return add2(mulf2(uv_of(shp, shp->elems[2 * ei]), ew.x), mulf2(uv_of(shp, shp->elems[2 * ei + 1]), ew.y));
    return add2(add2(mulf2(uv_of(shp, shp->elems[3 * ei]), ew.x), mulf2(uv_of(shp, shp->elems[3 * ei + 1]), ew.y)),
                mulf2(uv_of(shp, shp->elems[3 * ei + 2]), ew.z));
}

/* ---- scene.cpp: primitive tests ----------------------------------------------------------- */
/* scene.cpp:229-263 */
static int intersect_triangle(const ray3f* ray, vec3f v0, vec3f v1, vec3f v2, float* dist, vec4f* ew) {
    vec3f e1 = sub3(v1, v0);
    vec3f e2 = sub3(v2, v0);
    vec3f r = cross3(ray->d, e2);
    float den = dot3(r, e1);
    if (den == 0) return 0;
    float inv_den = 1.0f / den;
    vec3f c = sub3(ray->o, v0);
    float w1 = dot3(r, c) * inv_den;
    if (w1 < 0 || w1 > 1) return 0;
    vec3f s = cross3(c, e1);
    float w2 = dot3(s, ray->d) * inv_den;
    if (w2 < 0.0 || w1 + w2 > 1.0) return 0;
    float t = dot3(s, e2) * inv_den;
    if (t < ray->tmin || t > ray->tmax) return 0;
    *dist = t;
    ew->x = 1 - w1 - w2; ew->y = w1; ew->z = w2; ew->w = 0;
    return 1;
}
/* scene.cpp:267-281 */
static int intersect_point(const ray3f* ray, vec3f p, float r, float* dist, vec4f* ew) {
    vec3f w = sub3(p, ray->o);
    float t = dot3(w, ray->d) / dot3(ray->d, ray->d);
    if (t < ray->tmin || t > ray->tmax) return 0;
    vec3f rp = eval_ray(ray, t);
    vec3f prp = sub3(p, rp);
    if (dot3(prp, prp) > r * r) return 0;
    *dist = t;
    ew->x = 1; ew->y = 0; ew->z = 0; ew->w = 0;
    return 1;
}
/* scene.cpp:285-307 */
static int intersect_line(const ray3f* ray, vec3f v0, vec3f v1, float r0, float r1, float* dist, vec4f* ew) {
    vec3f u = ray->d, v = sub3(v1, v0), w = sub3(ray->o, v0);
    float a = dot3(u, u), b = dot3(u, v), c = dot3(v, v), d = dot3(u, w), e = dot3(v, w);
    float det = a * c - b * b;
    if (det == 0) return 0;
    float t = (b * e - c * d) / det, s = (a * e - b * d) / det;
    if (t < ray->tmin || t > ray->tmax) return 0;
    s = clampf_(s, (float)0, (float)1);
    vec3f p0 = eval_ray(ray, t);
    ray3f lr; lr.o = v0; lr.d = sub3(v1, v0); lr.tmin = ray_eps; lr.tmax = FLT_MAX;
    vec3f p1 = eval_ray(&lr, s);
    vec3f p01 = sub3(p0, p1);
    float r = r0 * (1 - s) + r1 * s;
    if (dot3(p01, p01) > r * r) return 0;
    *dist = t;
    ew->x = 1 - s; ew->y = s; ew->z = 0; ew->w = 0;
    return 1;
}

/* one element of a shape, by kind */
static int intersect_elem(const o_shape* shp, int ei, const ray3f* tray, float* dist, vec4f* ew) {
    if (shp->kind == YRT_TRIANGLES) {
        const int32_t* e = shp->elems + 3 * ei;
        return intersect_triangle(tray, pos_of(shp, e[0]), pos_of(shp, e[1]), pos_of(shp, e[2]), dist, ew);
    }
    if (shp->kind == YRT_LINES) {
        const int32_t* e = shp->elems + 2 * ei;
        return intersect_line(tray, pos_of(shp, e[0]), pos_of(shp, e[1]), shp->radius[e[0]], shp->radius[e[1]], dist, ew);
    }
    int e = shp->elems[ei];
    return intersect_point(tray, pos_of(shp, e), shp->radius[e], dist, ew);
}

/* scene.cpp:311-342, brute force over a shape */
static int intersect_ray_shape(const o_shape* shp, const ray3f* ray, float* dist, int* iei, vec4f* iew) {
    int hit = 0;
    ray3f tray = *ray;
    for (int ei = 0; ei < shp->n_elems; ei++) {
        if (!intersect_elem(shp, ei, &tray, dist, iew)) continue;
        *iei = ei;
        tray.tmax = *dist;
        hit = 1;
    }
    return hit;
}
/* scene.cpp:346-367, brute force over the scene */
static intersection3f intersect_scene(const oracle_scene* scn, const ray3f* ray) {
    intersection3f isec = {-1, -1, {0, 0, 0, 0}, 0};
    int hit = 0;
    ray3f tray = *ray;
    for (int i = 0; i < scn->n_instances; i++) {
        const o_instance* ist = &scn->instances[i];
        ray3f lray = transform_ray_inverse(&ist->frame, &tray);
        if (!intersect_ray_shape(&scn->shapes[ist->shp], &lray, &isec.dist, &isec.ei, &isec.ew)) continue;
        tray.tmax = isec.dist;
        isec.ist = i;
        hit = 1;
    }
    if (!hit) { intersection3f none = {-1, -1, {0, 0, 0, 0}, 0}; return none; }
    return isec;
}

/* scene.cpp:371-383 */
static inline int intersect_check_bbox(const ray3f* ray, const bbox3f* bbox) {
    vec3f invd = v3(1.0f / ray->d.x, 1.0f / ray->d.y, 1.0f / ray->d.z);
    vec3f t0 = mul3(sub3(bbox->min, ray->o), invd);
    vec3f t1 = mul3(sub3(bbox->max, ray->o), invd);
    float tmp;
    if (invd.x < 0) { tmp = t0.x; t0.x = t1.x; t1.x = tmp; }
    if (invd.y < 0) { tmp = t0.y; t0.y = t1.y; t1.y = tmp; }
    if (invd.z < 0) { tmp = t0.z; t0.z = t1.z; t1.z = tmp; }
    float tmin = maxf_(t0.z, maxf_(t0.y, maxf_(t0.x, ray->tmin)));
    float tmax = minf_(t1.z, minf_(t1.y, minf_(t1.x, ray->tmax)));
    tmax *= 1.00000024f;
    return tmin <= tmax;
}

/* scene.cpp:386-442 */
static int intersect_bvh_shape(const o_shape* shp, const ray3f* ray, int any, float* dist, int* ei, vec4f* ew) {
    int node_stack[64];
    int node_cur = 0;
    node_stack[node_cur++] = 0;
    const bvh_tree* bvh = &shp->bvh;
    ray3f tray = *ray;
    int hit = 0;
    while (node_cur) {
        bvh_node node = bvh->nodes[node_stack[--node_cur]];
        if (!intersect_check_bbox(&tray, &node.bbox)) continue;
        if (!node.isleaf) {
            for (uint32_t i = node.start; i < node.start + node.count; i++) node_stack[node_cur++] = (int)i;
        } else {
            for (uint32_t i = node.start; i < node.start + node.count; i++) {
                if (!intersect_elem(shp, bvh->leaf_prims[i], &tray, dist, ew)) continue;
                hit = 1;
                tray.tmax = *dist;
                *ei = bvh->leaf_prims[i];
                if (any) return 1;
            }
        }
    }
    return hit;
}
/* scene.cpp:446-479 */
static int intersect_bvh_scene(const oracle_scene* scn, const ray3f* ray, int any, float* dist, int* ist, int* ei, vec4f* ew) {
    int node_stack[64];
    int node_cur = 0;
    node_stack[node_cur++] = 0;
    const bvh_tree* bvh = &scn->bvh;
    ray3f tray = *ray;
    int hit = 0;
    while (node_cur) {
        bvh_node node = bvh->nodes[node_stack[--node_cur]];
        if (!intersect_check_bbox(&tray, &node.bbox)) continue;
        if (!node.isleaf) {
            for (uint32_t i = node.start; i < node.start + node.count; i++) node_stack[node_cur++] = (int)i;
        } else {
            for (uint32_t i = node.start; i < node.start + node.count; i++) {
                int is = bvh->leaf_prims[i];
                const o_instance* inst = &scn->instances[is];
                ray3f lray = transform_ray_inverse(&inst->frame, &tray);
                if (!intersect_bvh_shape(&scn->shapes[inst->shp], &lray, any, dist, ei, ew)) continue;
                tray.tmax = *dist;
                *ist = is;
                hit = 1;
                if (any) return hit;
            }
        }
    }
    return hit;
}
/* scene.cpp:483-494 */
static intersection3f intersect_first(const oracle_scene* scn, const ray3f* ray) {
    intersection3f isec = {-1, -1, {0, 0, 0, 0}, 0};
    if (scn->n_instances == 0 || !intersect_bvh_scene(scn, ray, 0, &isec.dist, &isec.ist, &isec.ei, &isec.ew)) {
        intersection3f none = {-1, -1, {0, 0, 0, 0}, 0};
        return none;
    }
    return isec;
}
static int intersect_any(const oracle_scene* scn, const ray3f* ray) {
    intersection3f isec = {-1, -1, {0, 0, 0, 0}, 0};
    if (scn->n_instances == 0) return 0;
    return intersect_bvh_scene(scn, ray, 1, &isec.dist, &isec.ist, &isec.ei, &isec.ew);
}

/* ---- scene.cpp: BVH build ------------------------------------------------------------------ */
typedef struct { bbox3f bbox; vec3f center; int pid; } bound_prim;   /* scene.cpp:509-513 */

static bbox3f expand_bbox_pr(bbox3f bbox, vec3f p, float r) {        /* scene.cpp:521-523 */
    bbox3f b = {sub3(p, v3(r, r, r)), add3(p, v3(r, r, r))};
    return expand_bbox_b(bbox, b);
}

/* std::partition as libstdc++ implements it for bidirectional iterators (what the compiled reference runs) */
static int partition_prims(bound_prim* v, int first, int last, int axis, float half) {
#define PRED(p) ((axis == 0 ? (p).center.x : (axis == 1 ? (p).center.y : (p).center.z)) < half)
    for (;;) {
        for (;;) {
            if (first == last) return first;
            if (PRED(v[first])) ++first; else break;
        }
        --last;
        for (;;) {
            if (first == last) return first;
            if (!PRED(v[last])) --last; else break;
        }
        bound_prim tmp = v[first]; v[first] = v[last]; v[last] = tmp;
        ++first;
    }
#undef PRED
}

/* scene.cpp:607-639 with equalnum == false (raytrace.cpp:278) */
static int split_prims(bound_prim* sorted_prim, int start, int end, int* axis, int* mid) {
    bbox3f centroid_bbox = invalid_bbox3f;
    for (int i = start; i < end; i++) centroid_bbox = expand_bbox_p(centroid_bbox, sorted_prim[i].center);
    vec3f size = sub3(centroid_bbox.max, centroid_bbox.min);
    if (size.x == 0 && size.y == 0 && size.z == 0) return 0;
    if (size.x >= size.y && size.x >= size.z) *axis = 0;
    else if (size.y >= size.x && size.y >= size.z) *axis = 1;
    else *axis = 2;
    vec3f half = divf3(add3(centroid_bbox.min, centroid_bbox.max), 2);
    float h = *axis == 0 ? half.x : (*axis == 1 ? half.y : half.z);
    *mid = partition_prims(sorted_prim, start, end, *axis, h);
    return 1;
}

static int push_node(bvh_tree* bvh) {
    if (bvh->n_nodes == bvh->cap_nodes) {
        bvh->cap_nodes = bvh->cap_nodes ? 2 * bvh->cap_nodes : 64;
        bvh->nodes = (bvh_node*)realloc(bvh->nodes, sizeof(bvh_node) * (size_t)bvh->cap_nodes);
    }
    memset(&bvh->nodes[bvh->n_nodes], 0, sizeof(bvh_node));
    return bvh->n_nodes++;
}

/* scene.cpp:572-603 */
static void make_node(bvh_tree* bvh, int nid, bound_prim* leaf_prims, int start, int end) {
    bbox3f bbox = invalid_bbox3f;
    for (int i = start; i < end; i++) bbox = expand_bbox_b(bbox, leaf_prims[i].bbox);
    bvh->nodes[nid].bbox = bbox;
    int split = 0, axis = -1, mid = -1;
    if (end - start > 4) split = split_prims(leaf_prims, start, end, &axis, &mid);
    if (split && (mid <= start || mid >= end)) split = 0;   /* reference: assert (compiled out) then unbounded recursion */
    if (!split) {
        bvh->nodes[nid].isleaf = 1;
        bvh->nodes[nid].start = (uint32_t)start;
        bvh->nodes[nid].count = (uint16_t)(end - start);
    } else {
        bvh->nodes[nid].isleaf = 0;
        bvh->nodes[nid].axis = (uint8_t)axis;
        int first = push_node(bvh);
        push_node(bvh);
        bvh->nodes[nid].start = (uint32_t)first;
        bvh->nodes[nid].count = 2;
        make_node(bvh, first, leaf_prims, start, mid);
        make_node(bvh, first + 1, leaf_prims, mid, end);
    }
}
/* scene.cpp:643-658 */
static void build_bvh_prims(bvh_tree* bvh, bound_prim* prims, int n) {
    memset(bvh, 0, sizeof(*bvh));
    push_node(bvh);
    make_node(bvh, 0, prims, 0, n);
    bvh->n_prims = n;
    bvh->leaf_prims = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) bvh->leaf_prims[i] = prims[i].pid;
}
/* scene.cpp:525-549 */
static void build_bvh_shape(o_shape* shp) {
    int n = shp->n_elems;
    bound_prim* bp = (bound_prim*)malloc(sizeof(bound_prim) * (size_t)(n > 0 ? n : 1));
    for (int ei = 0; ei < n; ei++) {
        bbox3f bbox = invalid_bbox3f;
        if (shp->kind == YRT_POINTS) {
            int e = shp->elems[ei];
            bbox = expand_bbox_pr(bbox, pos_of(shp, e), shp->radius[e]);
        } else if (shp->kind == YRT_LINES) {
            const int32_t* e = shp->elems + 2 * ei;
            bbox = expand_bbox_pr(bbox, pos_of(shp, e[0]), shp->radius[e[0]]);
            bbox = expand_bbox_pr(bbox, pos_of(shp, e[1]), shp->radius[e[1]]);
        } else {
            const int32_t* e = shp->elems + 3 * ei;
            bbox = expand_bbox_pr(bbox, pos_of(shp, e[0]), 0);
            bbox = expand_bbox_pr(bbox, pos_of(shp, e[1]), 0);
            bbox = expand_bbox_pr(bbox, pos_of(shp, e[2]), 0);
        }
        bp[ei].bbox = bbox;
        bp[ei].center = divf3(add3(bbox.min, bbox.max), 2.0f);
        bp[ei].pid = ei;
    }
    build_bvh_prims(&shp->bvh, bp, n);
    free(bp);
}
/* scene.cpp:554-565 */
static void build_bvh_scene(oracle_scene* scn) {
    for (int s = 0; s < scn->n_shapes; s++) build_bvh_shape(&scn->shapes[s]);
    int n = scn->n_instances;
    bound_prim* bp = (bound_prim*)malloc(sizeof(bound_prim) * (size_t)(n > 0 ? n : 1));
    for (int ii = 0; ii < n; ii++) {
        const o_instance* ist = &scn->instances[ii];
        bbox3f bbox = bbox_to_world(&ist->frame, scn->shapes[ist->shp].bvh.nodes[0].bbox);
        bp[ii].bbox = bbox;
        bp[ii].center = divf3(add3(bbox.min, bbox.max), 2.0f);
        bp[ii].pid = ii;
    }
    build_bvh_prims(&scn->bvh, bp, n);
    free(bp);
}

/* ---- raytrace.cpp --------------------------------------------------------------------------- */
/* raytrace.cpp:6-37 */
static ray3f eval_camera(const yrt_camera* cam, vec2f uv) {
    float u = uv.x, v = uv.y;
    const float* f = cam->frame;
    vec3f o = v3(f[9], f[10], f[11]);
    vec3f x = v3(f[0], f[1], f[2]);
    vec3f y = mulf3(v3(f[3], f[4], f[5]), -1);
    vec3f z = v3(f[6], f[7], f[8]);
    float h = 2.0f * cam->focus * tanf(cam->fovy / 2.0f);
    float w = h * cam->aspect;
    float focus = cam->focus;
    vec3f q;
    q.x = o.x + (u - 0.5f) * w * x.x + (v - 0.5f) * h * y.x - focus * z.x;
    q.y = o.y + (u - 0.5f) * w * x.y + (v - 0.5f) * h * y.y - focus * z.y;
    q.z = o.z + (u - 0.5f) * w * x.z + (v - 0.5f) * h * y.z - focus * z.z;
    ray3f r;
    r.o = o;
    r.d = normalize3(sub3(q, o));
    r.tmin = ray_eps;
    r.tmax = FLT_MAX;
    return r;
}
/* raytrace.cpp:39-56 */
static vec3f lookup_texture(const o_texture* txt, int i, int j, int srgb) {
    const uint8_t* px = txt->px + 4 * ((size_t)j * (size_t)txt->w + (size_t)i);
    float r = px[0], g = px[1], b = px[2];
    float gamma = 2.2f;
    if (!srgb) gamma = 1.0f;
    vec3f v;
    v.x = fminf(1.0f, powf(r / 255.0f, gamma));
    v.y = fminf(1.0f, powf(g / 255.0f, gamma));
    v.z = fminf(1.0f, powf(b / 255.0f, gamma));
    return v;
}
/* raytrace.cpp:58-86 */
static vec3f eval_texture(const o_texture* txt, vec2f texcoord, int srgb) {
    float u = texcoord.x, v = texcoord.y;
    float w = (float)txt->w, h = (float)txt->h;
    float s = (float)(fmod((double)u, 1.0) * (double)w);
    float t = (float)(fmod((double)v, 1.0) * (double)h);
    int i = (int)floor((double)s), j = (int)floor((double)t);
    int i1 = (int)fmod((double)(i + 1), (double)w), j1 = (int)fmod((double)(j + 1), (double)h);
    float wi = s - (float)i, wj = t - (float)j;
    /* the reference indexes out of bounds for negative coordinates (UB); wrap like the CUDA path */
    i = ((i % txt->w) + txt->w) % txt->w; j = ((j % txt->h) + txt->h) % txt->h;
    i1 = ((i1 % txt->w) + txt->w) % txt->w; j1 = ((j1 % txt->h) + txt->h) % txt->h;
    vec3f cij = mulf3(mulf3(lookup_texture(txt, i, j, srgb), (1 - wi)), (1 - wj));
    vec3f ci1j = mulf3(mulf3(lookup_texture(txt, i1, j, srgb), wi), (1 - wj));
    vec3f cij1 = mulf3(mulf3(lookup_texture(txt, i, j1, srgb), (1 - wi)), wj);
    vec3f ci1j1 = mulf3(mulf3(lookup_texture(txt, i1, j1, srgb), wi), wj);
    return add3(add3(add3(cij, ci1j), cij1), ci1j1);
}

typedef struct { int64_t first_primary, first_reflect, any, max_depth; int depth_cap; } ray_counter;

/* raytrace.cpp:88-211.  `lights` is all instances (raytrace.cpp:241) filtered inside the loop (:126). */
static vec4f shade(const oracle_scene* scn, vec3f amb, const ray3f* ray, int depth, ray_counter* rc) {
    vec4f black = {0.0f, 0.0f, 0.0f, 1.0f};
    if (depth == 0) rc->first_primary++; else rc->first_reflect++;
    if (depth + 1 > rc->max_depth) rc->max_depth = depth + 1;
    intersection3f inter = intersect_first(scn, ray);
    if (inter.ei < 0) return black;
    int ei = inter.ei;
    vec4f ew = inter.ew;
    const o_instance* ist = &scn->instances[inter.ist];
    const o_shape* shp = &scn->shapes[ist->shp];
    const o_material* mat = &scn->materials[ist->mat];
    vec3f n = transform_direction(&ist->frame, eval_norm_shape(shp, ei, ew));   /* scene.h:216 */
    vec3f p = transform_point(&ist->frame, eval_pos_shape(shp, ei, ew));        /* scene.h:210 */
    vec3f c = v3(0.0f, 0.0f, 0.0f);
    vec3f kd = mat->kd, ks = mat->ks;
    const o_texture* texkd = mat->kd_txt >= 0 ? &scn->textures[mat->kd_txt] : NULL;
    const o_texture* texks = mat->ks_txt >= 0 ? &scn->textures[mat->ks_txt] : NULL;
    vec2f uv = eval_texcoord_shape(shp, ei, ew);
    vec3f la = mul3(amb, kd);
    if (texkd) la = mul3(la, eval_texture(texkd, uv, 1));
    for (int li = 0; li < scn->n_instances; li++) {
        const o_instance* light = &scn->instances[li];
        vec3f ke = scn->materials[light->mat].ke;
        if (ke.x > 0.0f && ke.y > 0.0f && ke.z > 0.0f) {
            const o_shape* lshp = &scn->shapes[light->shp];
            vec3f L = transform_point(&light->frame, sub3(pos_of(lshp, 0), p));
            vec3f l = normalize3(L);
            float r = length3(L);
            ray3f sr; sr.o = p; sr.d = l; sr.tmin = 0.01f; sr.tmax = r - 0.01f;
            rc->any++;
            if (!intersect_any(scn, &sr)) {
                float rs = mat->rs;
                float ns = (rs) ? 2 / powf(rs, 4.0f) - 2 : 1e6f;
                vec3f v = normalize3(sub3(ray->o, p));
                vec3f h = normalize3(add3(v, l));
                kd = mat->kd;
                ks = mat->ks;
                if (texkd) kd = mul3(kd, eval_texture(texkd, uv, 1));
                if (texks) ks = mul3(ks, eval_texture(texks, uv, 1));
                vec3f ld = mul3(kd, divf3(ke, r * r));
                vec3f ls = mul3(ks, divf3(ke, r * r));
                if (shp->kind == YRT_LINES) {
                    float prodnl = dot3(n, l), prodnh = dot3(n, h);
                    if (prodnl < 0.0f) prodnl *= -1;
                    if (prodnh < 0.0f) prodnh *= -1;
                    float sinnl = sqrtf(1.0f - prodnl), sinnh = sqrtf(1.0f - prodnh);
                    ld = mulf3(ld, sinnl);
                    ls = mulf3(ls, powf(sinnh, ns));
                } else {
                    ld = mulf3(ld, maxf_(0.0f, dot3(n, l)));
                    ls = mulf3(ls, powf(maxf_(0.0f, dot3(n, h)), ns));
                }
                c = add3(c, add3(ld, ls));
            }
        }
    }
    vec3f kr = mat->kr;
    if ((kr.x > 0.0f || kr.y > 0.0f || kr.z > 0.0f) && (rc->depth_cap <= 0 || depth + 1 < rc->depth_cap)) {
        vec3f v = normalize3(sub3(ray->o, p));
        vec3f dr = sub3(mulf3(mulf3(n, 2.0f), dot3(n, v)), v);
        ray3f newr; newr.o = p; newr.d = dr; newr.tmin = ray_eps; newr.tmax = FLT_MAX;
        vec4f col = shade(scn, amb, &newr, depth + 1, rc);
        c = add3(c, v3(col.x * kr.x, col.y * kr.y, col.z * kr.z));
    }
    c = add3(c, la);
    vec4f L = {c.x, c.y, c.z, 1.0f};
    return L;
}

/* ---- public ---------------------------------------------------------------------------------- */
static void* dup_mem(const void* p, size_t n) {
    void* q = malloc(n ? n : 1);
    if (n && p) memcpy(q, p, n); else if (n) memset(q, 0, n);
    return q;
}

int oracle_scene_create(const yrt_scene_desc* d, oracle_scene** out) {
    if (!d || !out) return YRT_ERR_INVALID;
    oracle_scene* s = (oracle_scene*)calloc(1, sizeof(oracle_scene));
    s->n_shapes = d->n_shapes; s->n_instances = d->n_instances; s->n_materials = d->n_materials; s->n_textures = d->n_textures;
    s->elem_idx = (int32_t*)dup_mem(d->elem_idx, sizeof(int32_t) * (size_t)d->n_elem_idx);
    s->pos = (float*)dup_mem(d->pos, sizeof(float) * 3 * (size_t)d->n_verts);
    s->norm = (float*)dup_mem(d->norm, sizeof(float) * 3 * (size_t)d->n_verts);
    s->uv = (float*)dup_mem(d->uv, sizeof(float) * 2 * (size_t)d->n_verts);
    s->radius = (float*)dup_mem(d->radius, sizeof(float) * (size_t)d->n_verts);
    s->texels = (uint8_t*)dup_mem(d->tex_rgba8, (size_t)(d->n_textures ? d->tex_bytes : 0));
    s->shapes = (o_shape*)calloc((size_t)(d->n_shapes ? d->n_shapes : 1), sizeof(o_shape));
    for (int i = 0; i < d->n_shapes; i++) {
        o_shape* sh = &s->shapes[i];
        sh->kind = d->shape_kind[i];
        sh->n_elems = d->shape_elem_cnt[i];
        sh->n_verts = d->shape_vert_cnt[i];
        sh->has_uv = (d->uv != NULL) && (d->shape_has_uv ? d->shape_has_uv[i] != 0 : 1);
        sh->elems = s->elem_idx + d->shape_elem_off[i];
        sh->pos = s->pos + 3 * (size_t)d->shape_vert_off[i];
        sh->norm = s->norm + 3 * (size_t)d->shape_vert_off[i];
        sh->uv = s->uv + 2 * (size_t)d->shape_vert_off[i];
        sh->radius = s->radius + (size_t)d->shape_vert_off[i];
    }
    s->instances = (o_instance*)calloc((size_t)(d->n_instances ? d->n_instances : 1), sizeof(o_instance));
    for (int i = 0; i < d->n_instances; i++) {
        const float* f = d->inst_frame + 12 * (size_t)i;
        o_instance* in = &s->instances[i];
        in->frame.x = v3(f[0], f[1], f[2]); in->frame.y = v3(f[3], f[4], f[5]); in->frame.z = v3(f[6], f[7], f[8]); in->frame.o = v3(f[9], f[10], f[11]);
        in->shp = d->inst_shape[i];
        in->mat = d->inst_mat[i];
    }
    s->materials = (o_material*)calloc((size_t)(d->n_materials ? d->n_materials : 1), sizeof(o_material));
    for (int i = 0; i < d->n_materials; i++) {
        o_material* m = &s->materials[i];
        m->ke = v3(d->mat_ke[3 * i], d->mat_ke[3 * i + 1], d->mat_ke[3 * i + 2]);
        m->kd = v3(d->mat_kd[3 * i], d->mat_kd[3 * i + 1], d->mat_kd[3 * i + 2]);
        m->ks = v3(d->mat_ks[3 * i], d->mat_ks[3 * i + 1], d->mat_ks[3 * i + 2]);
        m->kr = v3(d->mat_kr[3 * i], d->mat_kr[3 * i + 1], d->mat_kr[3 * i + 2]);
        m->rs = d->mat_rs[i];
        m->kd_txt = d->mat_kd_tex[i] >= 0 ? d->mat_kd_tex[i] : -1;
        m->ks_txt = d->mat_ks_tex[i] >= 0 ? d->mat_ks_tex[i] : -1;
    }
    s->textures = (o_texture*)calloc((size_t)(d->n_textures ? d->n_textures : 1), sizeof(o_texture));
    for (int i = 0; i < d->n_textures; i++) {
        s->textures[i].w = d->tex_w[i];
        s->textures[i].h = d->tex_h[i];
        s->textures[i].px = s->texels + d->tex_off[i];
    }
    build_bvh_scene(s);
    *out = s;
    return YRT_OK;
}

void oracle_scene_destroy(oracle_scene* s) {
    if (!s) return;
    for (int i = 0; i < s->n_shapes; i++) { free(s->shapes[i].bvh.nodes); free(s->shapes[i].bvh.leaf_prims); }
    free(s->bvh.nodes); free(s->bvh.leaf_prims);
    free(s->shapes); free(s->instances); free(s->materials); free(s->textures);
    free(s->elem_idx); free(s->pos); free(s->norm); free(s->uv); free(s->radius); free(s->texels);
    free(s);
}

void oracle_scene_info(const oracle_scene* s, int64_t out[4]) {
    out[0] = s->bvh.n_nodes;
    out[1] = 0;
    for (int i = 0; i < s->n_shapes; i++) out[1] += s->shapes[i].bvh.n_nodes;
    out[2] = 0;
    for (int i = 0; i < s->n_instances; i++) {
        vec3f ke = s->materials[s->instances[i].mat].ke;
        if (ke.x > 0.0f && ke.y > 0.0f && ke.z > 0.0f) out[2]++;
    }
    out[3] = 0;
}

int oracle_image_width(const yrt_camera* cam, int resolution) { return (int)roundf(cam->aspect * (float)resolution); }   /* raytrace.cpp:216 */

int oracle_render_rows(const oracle_scene* scn, const yrt_camera* cam, const float amb_[3], int width, int height, int samples,
                       int max_depth, int n_threads, int row0, int row1, float* rgba, int64_t counts[4]) {
    if (!scn || !cam || !rgba || width <= 0 || height <= 0 || samples <= 0) return YRT_ERR_INVALID;
    vec3f amb = v3(amb_[0], amb_[1], amb_[2]);
    int64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (n_threads < 1) n_threads = 1;
    /* raytrace.cpp:228-251 (the reference loops i outer / j inner; pixels are independent) */
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : c0, c1, c2) reduction(max : c3)
    for (int j = row0; j < row1; j++) {
        ray_counter rc = {0, 0, 0, 0, max_depth};
        for (int i = 0; i < width; i++) {
            vec4f px = {0, 0, 0, 0};
            for (int jj = 0; jj < samples; jj++) {
                for (int ii = 0; ii < samples; ii++) {
                    vec2f uv = {(i + (ii + 0.5f) / samples) / width, (j + (jj + 0.5f) / samples) / height};
                    ray3f raggio = eval_camera(cam, uv);
                    vec4f s = shade(scn, amb, &raggio, 0, &rc);
                    px.x += s.x; px.y += s.y; px.z += s.z; px.w += s.w;
                }
            }
            float* o = rgba + 4 * ((size_t)j * (size_t)width + (size_t)i);
            o[0] = px.x / (float)(samples * samples);
            o[1] = px.y / (float)(samples * samples);
            o[2] = px.z / (float)(samples * samples);
            o[3] = 1.0f;
        }
        c0 += rc.first_primary; c1 += rc.first_reflect; c2 += rc.any;
        if (rc.max_depth > c3) c3 = rc.max_depth;
    }
    if (counts) { counts[0] = c0; counts[1] = c1; counts[2] = c2; counts[3] = c3; }
    return YRT_OK;
}

int oracle_render(const oracle_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples, int max_depth,
                  int n_threads, float* rgba, int64_t counts[4]) {
    return oracle_render_rows(scn, cam, amb, width, height, samples, max_depth, n_threads, 0, height, rgba, counts);
}

static void fill_ids(const oracle_scene* scn, const intersection3f* isec, int32_t* ids, float* dist, float* uv) {
    if (isec->ei >= 0) {
        ids[0] = isec->ist; ids[1] = scn->instances[isec->ist].shp; ids[2] = isec->ei;
        if (dist) *dist = isec->dist;
        if (uv) { uv[0] = isec->ew.y; uv[1] = isec->ew.z; }
    } else {
        ids[0] = ids[1] = ids[2] = -1;
        if (dist) *dist = 0.f;
        if (uv) { uv[0] = uv[1] = 0.f; }
    }
}

int oracle_trace_primary(const oracle_scene* scn, const yrt_camera* cam, int width, int height, int samples, int brute_force,
                         int32_t* ids, float* dist, float* uv) {
    if (!scn || !cam || !ids) return YRT_ERR_INVALID;
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < height; j++)
        for (int i = 0; i < width; i++)
            for (int jj = 0; jj < samples; jj++)
                for (int ii = 0; ii < samples; ii++) {
                    size_t r = (((size_t)j * width + i) * samples + jj) * samples + ii;
                    vec2f tuv = {(i + (ii + 0.5f) / samples) / width, (j + (jj + 0.5f) / samples) / height};
                    ray3f ray = eval_camera(cam, tuv);
                    intersection3f isec = brute_force ? intersect_scene(scn, &ray) : intersect_first(scn, &ray);
                    fill_ids(scn, &isec, ids + 3 * r, dist ? dist + r : NULL, uv ? uv + 2 * r : NULL);
                }
    return YRT_OK;
}

int oracle_intersect_first(const oracle_scene* scn, const float* rays, int64_t n, int32_t* ids, float* dist, float* uv) {
    if (!scn || (n > 0 && (!rays || !ids))) return YRT_ERR_INVALID;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n; r++) {
        const float* q = rays + 8 * r;
        ray3f ray; ray.o = v3(q[0], q[1], q[2]); ray.d = v3(q[3], q[4], q[5]); ray.tmin = q[6]; ray.tmax = q[7];
        intersection3f isec = intersect_first(scn, &ray);
        fill_ids(scn, &isec, ids + 3 * r, dist ? dist + r : NULL, uv ? uv + 2 * r : NULL);
    }
    return YRT_OK;
}

int oracle_intersect_any(const oracle_scene* scn, const float* rays, int64_t n, uint8_t* occ) {
    if (!scn || (n > 0 && (!rays || !occ))) return YRT_ERR_INVALID;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n; r++) {
        const float* q = rays + 8 * r;
        ray3f ray; ray.o = v3(q[0], q[1], q[2]); ray.d = v3(q[3], q[4], q[5]); ray.tmin = q[6]; ray.tmax = q[7];
        occ[r] = intersect_any(scn, &ray) ? 1 : 0;
    }
    return YRT_OK;
}

/* image.cpp:55-78 with exposure 0, use_filmic false, no_srgb false */
void oracle_tonemap(const float* in, int width, int height, uint8_t* out) {
    size_t n = (size_t)width * (size_t)height;
    for (size_t t = 0; t < n; t++) {
        float hx = in[4 * t] * powf(2, 0), hy = in[4 * t + 1] * powf(2, 0), hz = in[4 * t + 2] * powf(2, 0), hw = in[4 * t + 3];
        hx = powf(hx, 1 / 2.2f); hy = powf(hy, 1 / 2.2f); hz = powf(hz, 1 / 2.2f);
        out[4 * t] = (unsigned char)(clampf_(hx, 0.0f, 1.0f) * 255);
        out[4 * t + 1] = (unsigned char)(clampf_(hy, 0.0f, 1.0f) * 255);
        out[4 * t + 2] = (unsigned char)(clampf_(hz, 0.0f, 1.0f) * 255);
        out[4 * t + 3] = (unsigned char)(clampf_(hw, 0.0f, 1.0f) * 255);
    }
}
