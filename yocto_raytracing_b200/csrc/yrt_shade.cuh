// yrt_shade.cuh — attribute evaluation, texture lookup and the shade() arithmetic for one hit.
//
// Restates shade() (src/raytrace.cpp:88-211), eval_texture/lookup_texture (:39-86) and
// eval_pos/eval_norm/eval_texcoord (src/scene.h:159-219) on the flattened scene.  Operation order
// is the reference's; FMA contraction is off for this TU, so every value except the two powf()
// calls in the specular lobe (raytrace.cpp:174,179 — CUDA powf vs glibc powf, a few ulp, measured
// harmless in SURVEY §8a) is bit-identical to the reference on the same inputs.
#pragma once
#include "yrt_scene.cuh"

#ifndef YRT_SHADE_ONE_DIVIDE
#define YRT_SHADE_ONE_DIVIDE 1   /* 1: one divide instead of three for lights whose ke components are equal (exact) */
#endif

namespace yrt {

struct HitAttr {
    vec3 p;     // world position  (eval_pos(ist,…),  scene.h:210)
    vec3 n;     // world normal / line tangent (eval_norm(ist,…), scene.h:216)
    vec2 uv;    // eval_texcoord(shp,…), scene.h:193 ; (0,0) when the shape has no texcoords
    int mat;    // material index
    int kind;   // element kind of the shape
    int inst;   // instance index in scn->instances order
    int shape;  // shape index
    int ei;     // element index inside the shape
};

// world position only (what the shadow-ray stage needs)
YRT_HD vec3 eval_hit_pos(const SceneView& sv, int si, int prim, float w1, float w2, int& kind_out) {
    const float4* ir = sv.inst_recs + 4 * (size_t)si;
    float4 f0 = ld4(ir), f1 = ld4(ir + 1), f2 = ld4(ir + 2), f3 = ld4(ir + 3);
    frame3 f;
    f.x = xyz(f0); f.y = xyz(f1); f.z = xyz(f2); f.o = xyz(f3);
    int kind = ((unsigned)float_as_int(f3.w)) >> 28;
    kind_out = kind;
    const float4* pr = sv.prim_recs + 3 * (size_t)prim;
    float4 q0 = ld4(pr);
    vec3 lp;
    if (kind == 2) {
        lp = xyz(q0);                                   // scene.h:161
    } else if (kind == 1) {
        float4 q1 = ld4(pr + 1);
        float ewx = 1 - w1;                             // scene.cpp:304
        lp = xyz(q0) * ewx + xyz(q1) * w1;              // scene.h:163-164
    } else {
        const float4* ar = sv.prim_attrs + YRT_ATTR_STRIDE * (size_t)prim;
        float4 q1 = ld4(ar + 4), q2 = ld4(ar + 5);      // v1, v2 (the trace record holds the edges)
        float ewx = 1 - w1 - w2;                        // scene.cpp:260
        lp = xyz(q0) * ewx + xyz(q1) * w1 + xyz(q2) * w2;   // scene.h:166-168
    }
    return transform_point(f, lp);
}

YRT_HD void eval_hit(const SceneView& sv, int si, int prim, float w1, float w2, HitAttr& a) {
    const float4* ir = sv.inst_recs + 4 * (size_t)si;
    float4 f0 = ld4(ir), f1 = ld4(ir + 1), f2 = ld4(ir + 2), f3 = ld4(ir + 3);
    frame3 f;
    f.x = xyz(f0); f.y = xyz(f1); f.z = xyz(f2); f.o = xyz(f3);
    int sk = float_as_int(f3.w);
    a.kind = ((unsigned)sk) >> 28;
    a.shape = sk & 0x0fffffff;
    a.inst = float_as_int(f1.w);
    a.mat = float_as_int(f2.w);
    const float4* pr = sv.prim_recs + 3 * (size_t)prim;
    const float4* ar = sv.prim_attrs + YRT_ATTR_STRIDE * (size_t)prim;
    float4 q0 = ld4(pr), a0 = ld4(ar);
    a.ei = float_as_int(q0.w);
    vec3 lp, ln;
    if (a.kind == 2) {
        lp = xyz(q0);                                   // scene.h:161
        ln = xyz(a0);                                   // scene.h:178 (not normalised here)
        a.uv.x = a0.w;                                  // scene.h:195
        a.uv.y = ld4(ar + 1).w;
    } else if (a.kind == 1) {
        float4 q1 = ld4(pr + 1), a1 = ld4(ar + 1), a2 = ld4(ar + 2), a3 = ld4(ar + 3);
        float ewx = 1 - w1;
        lp = xyz(q0) * ewx + xyz(q1) * w1;
        ln = normalize(xyz(a0) * ewx + xyz(a1) * w1);   // scene.h:180-181
        vec2 t0, t1;
        t0.x = a0.w; t0.y = a1.w; t1.x = a2.w; t1.y = a3.x;
        a.uv = t0 * ewx + t1 * w1;                      // scene.h:197-198
    } else {
        float4 q1 = ld4(ar + 4), q2 = ld4(ar + 5), a1 = ld4(ar + 1), a2 = ld4(ar + 2), a3 = ld4(ar + 3);   // q1, q2 = v1, v2
        float ewx = 1 - w1 - w2;
        lp = xyz(q0) * ewx + xyz(q1) * w1 + xyz(q2) * w2;
        ln = normalize(xyz(a0) * ewx + xyz(a1) * w1 + xyz(a2) * w2);   // scene.h:183-185
        vec2 t0, t1, t2;
        t0.x = a0.w; t0.y = a1.w; t1.x = a2.w; t1.y = a3.x; t2.x = a3.y; t2.y = a3.z;
        a.uv = t0 * ewx + t1 * w1 + t2 * w2;            // scene.h:200-202
    }
    a.p = transform_point(f, lp);
    a.n = transform_direction(f, ln);
}

// lookup_texture with srgb=true (raytrace.cpp:39-56): the gamma decode of a byte is a pure
// function of the byte, so it is a 256-entry table filled by the host libm.
YRT_HD vec3 lookup_texture(const uint8_t* texels, int w, int i, int j, const float* lut) {
    const uint8_t* t = texels + 4 * ((size_t)j * (size_t)w + (size_t)i);
#if defined(__CUDA_ARCH__)
    uchar4 c = __ldg((const uchar4*)t);
    return mk3(lut[c.x], lut[c.y], lut[c.z]);
#else
    return mk3(lut[t[0]], lut[t[1]], lut[t[2]]);
#endif
}

// eval_texture (raytrace.cpp:58-86). fmod(u,1) is exact; its double product with the float
// width rounds once to float, i.e. equals the float product. Negative coordinates index out of
// bounds in the reference (UB); here indices are wrapped into range instead.
YRT_HD vec3 eval_texture(const SceneView& sv, int tex, const vec2& texcoord, const float* lut) {
    int4 info = sv.tex_info[tex];
    const uint8_t* texels = sv.tex_rgba8 + (((size_t)(unsigned)info.w << 32) | (size_t)(unsigned)info.z);
    int wi_ = info.x, hi_ = info.y;
    float w = (float)wi_, h = (float)hi_;
    float u = texcoord.x, v = texcoord.y;
    float s = (u - truncf(u)) * w;     // fmod(u, 1) * w
    float t = (v - truncf(v)) * h;
    int i = (int)floorf(s), j = (int)floorf(t);
    float fi = s - (float)i, fj = t - (float)j;   // wi, wj (raytrace.cpp:75-76)
    int i1 = (i + 1) % wi_, j1 = (j + 1) % hi_;   // fmod((i+1), w)
    // defensive wrap (no effect for u,v >= 0)
    i = ((i % wi_) + wi_) % wi_;   j = ((j % hi_) + hi_) % hi_;
    i1 = ((i1 % wi_) + wi_) % wi_; j1 = ((j1 % hi_) + hi_) % hi_;
    vec3 cij = lookup_texture(texels, wi_, i, j, lut) * (1 - fi) * (1 - fj);
    vec3 ci1j = lookup_texture(texels, wi_, i1, j, lut) * fi * (1 - fj);
    vec3 cij1 = lookup_texture(texels, wi_, i, j1, lut) * (1 - fi) * fj;
    vec3 ci1j1 = lookup_texture(texels, wi_, i1, j1, lut) * fi * fj;
    return cij + ci1j + cij1 + ci1j1;
}

struct Material {
    vec3 kd, ks, kr, ke;
    float ns;
    int kd_tex, ks_tex;
};

YRT_HD Material load_material(const SceneView& sv, int mat) {
    const float4* m = sv.mat_recs + 4 * (size_t)mat;
    float4 m0 = ld4(m), m1 = ld4(m + 1), m2 = ld4(m + 2), m3 = ld4(m + 3);
    Material r;
    r.kd = xyz(m0); r.ns = m0.w;
    r.ks = xyz(m1); r.kd_tex = float_as_int(m1.w);
    r.kr = xyz(m2); r.ks_tex = float_as_int(m2.w);
    r.ke = xyz(m3);
    return r;
}

// the light vector of raytrace.cpp:129-130 and the shadow ray of :131
YRT_HD void light_vector(const SceneView& sv, int k, const vec3& p, vec3& l, float& r, vec3& ke) {
    const float4* lr = sv.light_recs + 5 * (size_t)k;
    float4 l0 = ld4(lr), l1 = ld4(lr + 1), l2 = ld4(lr + 2), l3 = ld4(lr + 3), l4 = ld4(lr + 4);
    frame3 f;
    f.x = xyz(l0); f.y = xyz(l1); f.z = xyz(l2); f.o = xyz(l3);
    ke = mk3(l0.w, l1.w, l2.w);
    vec3 L = transform_point(f, xyz(l4) - p);
    l = normalize(L);
    r = length(L);
}

YRT_HD ray3 shadow_ray(const vec3& p, const vec3& l, float r) {
    ray3 sr;
    sr.o = p; sr.d = l; sr.tmin = 0.01f; sr.tmax = r - 0.01f;   // raytrace.cpp:131
    return sr;
}

// The light loop of shade() (raytrace.cpp:121-185) for a hit whose shadow rays have been traced.
// vis(k) says whether light k is unoccluded.  Returns c (sum over lights, in light order) and la.
template <class VisFn>
YRT_HD void shade_lights(const SceneView& sv, const HitAttr& a, const Material& m, const vec3& ray_o,
                         const vec3& amb, const float* lut, VisFn vis, vec3& c_out, vec3& la_out) {
    vec3 c = mk3(0.f, 0.f, 0.f);
    vec3 la = amb * m.kd;                                    // raytrace.cpp:116
    vec3 tkd = mk3(1.f, 1.f, 1.f), tks = mk3(1.f, 1.f, 1.f);
    if (m.kd_tex >= 0) {
        tkd = eval_texture(sv, m.kd_tex, a.uv, lut);
        la = la * tkd;                                       // raytrace.cpp:119
    }
    if (m.ks_tex >= 0) tks = eval_texture(sv, m.ks_tex, a.uv, lut);
    for (int k = 0; k < sv.n_lights; k++) {
        if (!vis(k)) continue;                               // raytrace.cpp:133
        vec3 l, ke;
        float r;
        light_vector(sv, k, a.p, l, r, ke);
        vec3 kd = m.kd, ks = m.ks;
        if (m.kd_tex >= 0) kd = kd * tkd;                    // raytrace.cpp:153-157
        if (m.ks_tex >= 0) ks = ks * tks;
#if YRT_SHADE_ONE_DIVIDE
        // ke / (r*r) is three IEEE divides by the same number (raytrace.cpp:159-160); a white light (ke.x, ke.y, ke.z the
        // same bits) needs one of them — same quotient, bit for bit
        const float rr = r * r;
        vec3 kel;
        if (float_as_int(ke.x) == float_as_int(ke.y) && float_as_int(ke.y) == float_as_int(ke.z)) { float q = ke.x / rr; kel = mk3(q, q, q); }
        else kel = ke / rr;
        vec3 ld = kd * kel;
        vec3 ls = ks * kel;
#else
        vec3 ld = kd * (ke / (r * r));                       // raytrace.cpp:159-160
        vec3 ls = ks * (ke / (r * r));
#endif
        // ks == (0,0,0) exactly (most materials of the instance scenes): ls = +0 * (finite lobe) = +0 and ld + 0 == ld,
        // so the half vector and the powf are skipped.  The lobe is finite whenever ld is: its base n.h (or
        // sqrt(1-|n.h|)) of finite unit vectors lies in [0, 1 + 3e-7] and ns <= 1e6; and if n, l or v is not finite,
        // ld is NaN already and so is the sum, with or without the lobe.
        const bool no_spec = ks.x == 0.0f && ks.y == 0.0f && ks.z == 0.0f && m.ns > 0.0f && m.ns <= 1e6f;
        if (a.kind == 1) {                                   // shp->lines.size() > 0, raytrace.cpp:162-175
            float prodnl = dot(a.n, l);
            if (prodnl < 0.0f) prodnl = -prodnl;
            ld = ld * sqrtf(1.0f - prodnl);
            if (!no_spec) {
                vec3 v = normalize(ray_o - a.p);             // raytrace.cpp:147
                vec3 h = normalize(v + l);
                float prodnh = dot(a.n, h);
                if (prodnh < 0.0f) prodnh = -prodnh;
                ls = ls * powf(sqrtf(1.0f - prodnh), m.ns);
            }
        } else {                                             // raytrace.cpp:176-180
            ld = ld * rmax(0.0f, dot(a.n, l));
            if (!no_spec) {
                vec3 v = normalize(ray_o - a.p);             // raytrace.cpp:147
                vec3 h = normalize(v + l);
                ls = ls * powf(rmax(0.0f, dot(a.n, h)), m.ns);
            }
        }
        c = c + (ld + ls);                                   // raytrace.cpp:182
    }
    c_out = c;
    la_out = la;
}

YRT_HD bool is_reflective(const Material& m) { return m.kr.x > 0.0f || m.kr.y > 0.0f || m.kr.z > 0.0f; }   // :190

// the mirror ray of raytrace.cpp:192-199 (direction is NOT re-normalised)
YRT_HD ray3 reflection_ray(const HitAttr& a, const vec3& ray_o) {
    vec3 v = normalize(ray_o - a.p);
    vec3 dr = (a.n * 2.0f * dot(a.n, v)) - v;
    ray3 r;
    r.o = a.p; r.d = dr; r.tmin = YRT_RAY_EPS; r.tmax = FLT_MAX;
    return r;
}

// close one recursion level (raytrace.cpp:203,206): c + col*kr, then + la
YRT_HD vec3 combine_reflection(const vec3& c, const vec3& col, const vec3& kr, const vec3& la) {
    vec3 r = c + mk3(col.x * kr.x, col.y * kr.y, col.z * kr.z);
    return r + la;
}

// One shade() invocation for a ray that hit something (raytrace.cpp:95-210 without the recursion):
// returns true when a mirror ray must be traced first (then c, kr, la are the values the caller
// keeps until the recursion returns and `refl` is the ray), false when `value` is final (= c + la).
template <class VisFn>
YRT_HD bool shade_hit(const SceneView& sv, int si, int prim, float w1, float w2, const vec3& ray_o, const vec3& amb,
                      const float* lut, VisFn vis, bool allow_reflection, vec3& value, vec3& c, vec3& kr, vec3& la,
                      ray3& refl) {
    HitAttr at;
    eval_hit(sv, si, prim, w1, w2, at);
    Material m = load_material(sv, at.mat);
    shade_lights(sv, at, m, ray_o, amb, lut, vis, c, la);
    if (is_reflective(m) && allow_reflection) {
        kr = m.kr;
        refl = reflection_ray(at, ray_o);
        return true;
    }
    value = c + la;   // raytrace.cpp:206
    return false;
}

}  // namespace yrt
