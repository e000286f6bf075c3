// yrt_math.cuh — op-for-op restatement of the reference's float32 math on the hot path.
//
// Everything here is __host__ __device__ so the very same code is (a) the device path and
// (b) unit-tested on the CPU against the compiled reference (tests/, tools/host_emu).
// Bit-exactness contract (SURVEY.md §8a "Arithmetic contract"):
//   * + - * are never contracted into FMA: this TU is compiled with nvcc -fmad=false, and the
//     host build targets baseline x86-64 (no FMA instructions exist there);
//   * / and sqrtf are IEEE-rounded (nvcc defaults -prec-div=true -prec-sqrt=true -ftz=false);
//   * min/max/clamp are the reference's ternaries (src/vmath.h:215-217), NOT fminf/fmaxf.
// Citations are to /root/reference/src.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define YRT_HD __host__ __device__ __forceinline__
#else
#define YRT_HD inline
#endif

namespace yrt {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };

// ---- src/vmath.h:67-116 ------------------------------------------------------------------
YRT_HD vec3 mk3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
YRT_HD vec3 operator-(const vec3& a) { return mk3(-a.x, -a.y, -a.z); }
YRT_HD vec3 operator+(const vec3& a, const vec3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
YRT_HD vec3 operator-(const vec3& a, const vec3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
YRT_HD vec3 operator*(const vec3& a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
YRT_HD vec3 operator/(const vec3& a, float b) { return mk3(a.x / b, a.y / b, a.z / b); }   // vmath.h:81 true divides
YRT_HD vec3 operator*(const vec3& a, const vec3& b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
YRT_HD vec2 operator+(const vec2& a, const vec2& b) { vec2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
YRT_HD vec2 operator*(const vec2& a, float b) { vec2 r; r.x = a.x * b; r.y = a.y * b; return r; }

// vmath.h:112  (a.x*b.x + a.y*b.y) + a.z*b.z, left to right, no contraction
YRT_HD float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vmath.h:116
YRT_HD float length(const vec3& a) { return sqrtf(dot(a, a)); }
// vmath.h:118-122: one divide, three multiplies; zero-length input returned unchanged
YRT_HD vec3 normalize(const vec3& a) {
    float l = length(a);
    if (l == 0) return a;
    return a * (1 / l);
}
// vmath.h:124-127
YRT_HD vec3 cross(const vec3& a, const vec3& b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// vmath.h:215-217 (ternaries: a NaN first operand is dropped, a NaN second operand is kept)
YRT_HD float rmin(float x, float y) { return (x < y) ? x : y; }
YRT_HD float rmax(float x, float y) { return (x > y) ? x : y; }
YRT_HD float rclamp(float x, float a, float b) { return rmin(rmax(x, a), b); }

// ---- frames: src/vmath.h:145-175 -----------------------------------------------------------
struct frame3 { vec3 x, y, z, o; };

// vmath.h:152  ((x*b.x + y*b.y) + z*b.z) + o
YRT_HD vec3 transform_point(const frame3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.o; }
// vmath.h:161
YRT_HD vec3 transform_vector(const frame3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vmath.h:169
YRT_HD vec3 transform_direction(const frame3& a, const vec3& b) { return normalize(transform_vector(a, b)); }
// vmath.h:156-159
YRT_HD vec3 transform_point_inverse(const frame3& a, const vec3& b) {
    vec3 bo = b - a.o;
    return mk3(dot(a.x, bo), dot(a.y, bo), dot(a.z, bo));
}
// vmath.h:165-175
YRT_HD vec3 transform_direction_inverse(const frame3& a, const vec3& b) {
    return normalize(mk3(dot(a.x, b), dot(a.y, b), dot(a.z, b)));
}

// ---- rays: src/vmath.h:264-278 -------------------------------------------------------------
#define YRT_RAY_EPS 1e-4f   /* vmath.h:264 */
struct ray3 { vec3 o, d; float tmin, tmax; };
YRT_HD vec3 eval_ray(const ray3& r, float t) { return r.o + r.d * t; }   // vmath.h:273
// vmath.h:275-278: tmin/tmax are copied, the direction is re-normalised
YRT_HD ray3 transform_ray_inverse(const frame3& f, const ray3& r) {
    ray3 q;
    q.o = transform_point_inverse(f, r.o);
    q.d = transform_direction_inverse(f, r.d);
    q.tmin = r.tmin;
    q.tmax = r.tmax;
    return q;
}

// ---- primitive tests: src/scene.cpp:229-307 ------------------------------------------------
// Möller–Trumbore, two-sided, non-watertight (scene.cpp:229-263). On a hit writes t and the
// two barycentrics (ew = {1-w1-w2, w1, w2, 0}).
// The two edges e1 = v1 - v0, e2 = v2 - v0 (scene.cpp:236-237) do not depend on the ray: the build stores them (the same
// two float subtractions, done once) and the traversal calls this form.
YRT_HD bool intersect_triangle_edges(const ray3& ray, const vec3& v0, const vec3& e1, const vec3& e2, float& dist,
                                     float& ow1, float& ow2) {
    vec3 r = cross(ray.d, e2);
    float den = dot(r, e1);
    if (den == 0) return false;
    float inv_den = 1.0f / den;
    vec3 c = ray.o - v0;
    float w1 = dot(r, c) * inv_den;
    if (w1 < 0 || w1 > 1) return false;
    vec3 s = cross(c, e1);
    float w2 = dot(s, ray.d) * inv_den;
    // scene.cpp:252 compares in double (`w2 < 0.0 || w1 + w2 > 1.0`): the float sum w1+w2 is
    // rounded to float first (FLT_EVAL_METHOD 0) and only then widened, so float compares agree.
    if (w2 < 0.0f || w1 + w2 > 1.0f) return false;
    float t = dot(s, e2) * inv_den;
    if (t < ray.tmin || t > ray.tmax) return false;
    dist = t;
    ow1 = w1;
    ow2 = w2;
    return true;
}

YRT_HD bool intersect_triangle(const ray3& ray, const vec3& v0, const vec3& v1, const vec3& v2, float& dist,
                               float& ow1, float& ow2) {
    return intersect_triangle_edges(ray, v0, v1 - v0, v2 - v0, dist, ow1, ow2);
}

// scene.cpp:267-281; ew = {1,0,0,0}
YRT_HD bool intersect_point(const ray3& ray, const vec3& p, float r, float& dist) {
    vec3 w = p - ray.o;
    float t = dot(w, ray.d) / dot(ray.d, ray.d);
    if (t < ray.tmin || t > ray.tmax) return false;
    vec3 rp = eval_ray(ray, t);
    vec3 prp = p - rp;
    if (dot(prp, prp) > r * r) return false;
    dist = t;
    return true;
}

// scene.cpp:285-307; ew = {1-s, s, 0, 0}. The t range is checked BEFORE s is clamped and
// dist is the parameter of closest approach, not of the cylinder surface.
YRT_HD bool intersect_line(const ray3& ray, const vec3& v0, const vec3& v1, float r0, float r1, float& dist,
                           float& os) {
    vec3 u = ray.d, v = v1 - v0, w = ray.o - v0;
    float a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w);
    float det = a * c - b * b;
    if (det == 0) return false;
    float t = (b * e - c * d) / det, s = (a * e - b * d) / det;
    if (t < ray.tmin || t > ray.tmax) return false;
    s = rclamp(s, 0.0f, 1.0f);
    vec3 p0 = eval_ray(ray, t);
    vec3 p1 = v0 + (v1 - v0) * s;   // eval_ray(ray3f{v0, v1 - v0}, s)
    vec3 p01 = p0 - p1;
    float r = r0 * (1 - s) + r1 * s;
    if (dot(p01, p01) > r * r) return false;
    dist = t;
    os = s;
    return true;
}

// ---- slab test: src/scene.cpp:371-383 ------------------------------------------------------
// Same accept decision as the reference, restated for issue rate (the traversal is ALU-issue bound):
//   * invd = 1/d and the three `invd < 0` swap predicates are hoisted out (computed once per
//     (ray, space) instead of once per call; same IEEE divides);
//   * the swap is a select of which product feeds t0/t1;
//   * the ternary max/min chains become FMNMX: max(a, b) = (a > b) ? a : b returns b when a is NaN,
//     exactly like fmaxf(a, b) with a non-NaN b — and in scene.cpp:378-379 only the FIRST operand
//     (a slab distance, NaN when 0 * inf) can be NaN, the running value never is.  (+0 vs -0 may
//     differ, which no comparison below can see.)
// Returns the accept decision and the entry distance (for near/far ordering only).
struct raysigns { bool x, y, z; };
YRT_HD raysigns signs_of(const vec3& invd) { raysigns s; s.x = invd.x < 0; s.y = invd.y < 0; s.z = invd.z < 0; return s; }

YRT_HD bool intersect_check_bbox(const vec3& o, const vec3& invd, const raysigns& sg, float rtmin, float rtmax, float lox,
                                 float loy, float loz, float hix, float hiy, float hiz, float& tenter) {
    float ax = (lox - o.x) * invd.x, bx = (hix - o.x) * invd.x;
    float ay = (loy - o.y) * invd.y, by = (hiy - o.y) * invd.y;
    float az = (loz - o.z) * invd.z, bz = (hiz - o.z) * invd.z;
    float t0x = sg.x ? bx : ax, t1x = sg.x ? ax : bx;
    float t0y = sg.y ? by : ay, t1y = sg.y ? ay : by;
    float t0z = sg.z ? bz : az, t1z = sg.z ? az : bz;
    float tmin = fmaxf(t0z, fmaxf(t0y, fmaxf(t0x, rtmin)));
    float tmax = fminf(t1z, fminf(t1y, fminf(t1x, rtmax)));
    tmax *= 1.00000024f;
    tenter = tmin;
    return tmin <= tmax;
}

// Slab test used by the traversal.  Nodes store each child box as centre c and half-extent h' (inflated at build
// time, see box_center_half):  t0 = (c-o)*invd - h'*|invd|,  t1 = (c-o)*invd + h'*|invd|  needs NO per-axis near/far
// select and is 9 FFMA per box on the FMA pipe; only the two max3/min3 reductions and the compare stay on the ALU
// pipe.  The first version (FADD+FMUL, FSETP+FSEL ternaries, 6 selects per box) was ALU-pipe bound: ALU 70 %, FMA
// 27 % busy (profiles/).
// Conservative by construction.  With u = 2^-24: tc = fl(c*invd - fl(o*invd)) is off by at most u(|o invd| + |tc|) and
// |tc| <= |t| + h'|invd|; the second FFMA adds u|t|.  h' = h(1+16u) exceeds the true half-extent h by more than the
// 4u h it needs to (a) cover the rounding of c and h themselves and (b) absorb the u h'|invd| term: every computed near
// (far) distance is at most u M + 2u|t| above (below) the exact slab distance of a box that still CONTAINS the true
// box, M = max_a |o_a invd_a| per ray and space.  The reference accepts tmin <= tmax(1+4u) on values carrying 2u
// relative error, i.e. exact T0 <= T1(1+9u); chaining the bounds,
//       tmin <= tmax (1+16u) + 8u M
// accepts a SUPERSET of what scene.cpp:371-383 accepts for the true box: a box the reference would enter is never
// culled, and what is found inside is decided by the exact primitive tests, so hits are unchanged
// (tests/test_host_emu.py audits every box test against the reference's formula: 0 false rejects).
// NaN (inf-inf for direction components that are exactly 0) is dropped by FMNMX like in the reference's ternaries:
// that axis does not constrain (|invd| = inf keeps h'|invd| = inf or NaN, never a finite bound).
#ifndef YRT_APPROX_RCP
#define YRT_APPROX_RCP 1   /* slab rays use MUFU.RCP (1 ulp) instead of the correctly rounded 1/d, see rcp_slab in yrt_trace.cuh */
#endif
// With YRT_APPROX_RCP each invd_a is off by a further 2u relative, i.e. every slab distance by 2u|t| more: the chain
// above becomes T0 <= T1 (1+9u)(1+4u)/(1-4u) ~ T1 (1+17u), so the accept factor grows from 1+16u to 1+24u; the inflated
// half-extents keep their margin ((1+16u)(1-2u) > 1+13u against the 4u they need).
#if YRT_APPROX_RCP
#define YRT_SLAB_ACCEPT 1.0000014305114746f   /* 1 + 24u */
#else
#define YRT_SLAB_ACCEPT 1.00000095f           /* 1 + 16u */
#endif
struct slabray { vec3 invd, noi, ainv; float pad; };

YRT_HD slabray make_slabray(const vec3& o, const vec3& invd) {
    slabray r;
    r.invd = invd;
    r.noi = mk3(-(o.x * invd.x), -(o.y * invd.y), -(o.z * invd.z));
    r.ainv = mk3(fabsf(invd.x), fabsf(invd.y), fabsf(invd.z));
    float m = fmaxf(fmaxf(fmaxf(fabsf(o.x) * r.ainv.x, fabsf(o.y) * r.ainv.y), fabsf(o.z) * r.ainv.z), 0.0f);
    r.pad = m * 4.7683716e-7f;                                              // 8u M
    return r;
}

YRT_HD bool slab_test_ch(const slabray& r, float rtmin, float rtmax, float cx, float cy, float cz, float hx, float hy,
                         float hz, float& tenter) {
    float tcx = fmaf(cx, r.invd.x, r.noi.x), tcy = fmaf(cy, r.invd.y, r.noi.y), tcz = fmaf(cz, r.invd.z, r.noi.z);
    float t0x = fmaf(hx, -r.ainv.x, tcx), t1x = fmaf(hx, r.ainv.x, tcx);
    float t0y = fmaf(hy, -r.ainv.y, tcy), t1y = fmaf(hy, r.ainv.y, tcy);
    float t0z = fmaf(hz, -r.ainv.z, tcz), t1z = fmaf(hz, r.ainv.z, tcz);
    float tmin = fmaxf(t0z, fmaxf(t0y, fmaxf(t0x, rtmin)));
    float tmax = fminf(t1z, fminf(t1y, fminf(t1x, rtmax)));
    tenter = tmin;
    return tmin <= fmaf(tmax, YRT_SLAB_ACCEPT, r.pad);
}

// centre / half-extent of a box; the half-extent is inflated by 1 + 16u: [c-h', c+h'] contains [lo, hi] with the
// margin the slab test's error analysis needs (see slab_test_ch)
YRT_HD void box_center_half(float lo, float hi, float& c, float& h) {
    c = (lo + hi) * 0.5f;
    float d = fmaxf(hi - c, c - lo);
    h = d * 1.00000095f;
    if (!(h >= 0.f)) h = 0.f;
}

// ---- camera: src/raytrace.cpp:6-37 ---------------------------------------------------------
// h = 2*focus*tanf(fovy/2) and w = h*aspect are computed once on the host (raytrace.cpp:21-22).
struct camera_k { frame3 frame; float w, h, focus; };

YRT_HD ray3 eval_camera(const camera_k& cam, float u, float v) {
    vec3 o = cam.frame.o;
    vec3 x = cam.frame.x;
    vec3 y = cam.frame.y * -1;   // raytrace.cpp:18 (vec3f * float; int -1 promoted)
    vec3 z = cam.frame.z;
    float w = cam.w, h = cam.h, focus = cam.focus;
    vec3 q;
    // raytrace.cpp:28-30, left to right: ((o + ((u-.5)*w)*x) + ((v-.5)*h)*y) - focus*z
    q.x = o.x + (u - 0.5f) * w * x.x + (v - 0.5f) * h * y.x - focus * z.x;
    q.y = o.y + (u - 0.5f) * w * x.y + (v - 0.5f) * h * y.y - focus * z.y;
    q.z = o.z + (u - 0.5f) * w * x.z + (v - 0.5f) * h * y.z - focus * z.z;
    ray3 r;
    r.o = o;
    r.d = normalize(q - o);
    r.tmin = YRT_RAY_EPS;   // vmath.h:269-270 defaults
    r.tmax = FLT_MAX;
    return r;
}

// sample (ii,jj) of pixel (i,j): raytrace.cpp:236-239, int -> float promotions as written
YRT_HD void sample_uv(int i, int j, int ii, int jj, int samples, int width, int height, float& u, float& v) {
    u = ((float)i + ((float)ii + 0.5f) / (float)samples) / (float)width;
    v = ((float)j + ((float)jj + 0.5f) / (float)samples) / (float)height;
}

}  // namespace yrt
