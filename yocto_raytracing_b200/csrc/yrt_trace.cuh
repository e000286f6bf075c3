// yrt_trace.cuh — two-level BVH traversal (closest-hit and any-hit) for one ray.
//
// Replaces intersect_bvh(scene…) (src/scene.cpp:446-479) and intersect_bvh(shape…)
// (src/scene.cpp:386-442).  What is kept bit-for-bit: the ray handed to the primitive tests
// (transform_ray_inverse per instance, tmin/tmax copied, src/vmath.h:275-278), the primitive
// tests themselves, the accept rule `t > tmax` rejects (so equal t is accepted and tmax shrinks
// monotonically, scene.cpp:256,271,295) and the slab test's accept decision (scene.cpp:371-383).
// What is free (SURVEY finding 4): tree topology and visit order — here an LBVH with near-child
// first ordering.  Exact-distance ties (frequent in instance10000: overlapping instances with
// coplanar faces) are resolved like the reference does, by its visit order, through rank tables.
//
// The same code runs on the device (kernels in yrt_render.cu) and on the host (tools/host_emu,
// tests only).
#pragma once
#include "yrt_scene.cuh"

#ifndef YRT_VISIT_V2
#define YRT_VISIT_V2 1
#endif
#ifndef YRT_ANY_UNORDERED
#define YRT_ANY_UNORDERED 1  /* any-hit rays: skip the near/far ordering of the two children (the answer is order independent; -6 % kernel time) */
#endif

#ifndef YRT_STACK_TOP_REG
#define YRT_STACK_TOP_REG 0   /* 1: the top of the traversal stack lives in a register (a pop hands out the next node without waiting for a load) */
#endif
#ifndef YRT_PREFETCH_PUSH
#define YRT_PREFETCH_PUSH 0   /* 1: a postponed child node is prefetched into L1 when it is pushed */
#endif
#ifndef YRT_RESTORE_OD
#define YRT_RESTORE_OD 0   /* 1: an instance exit also restores the ray's origin and direction (the first version; dead values) */
#endif
#ifndef YRT_WORLD_SMEM
#define YRT_WORLD_SMEM 0   /* 1: the world-space ray waits in shared memory while the lane is inside an instance (13 registers less) */
#endif
#if YRT_WORLD_SMEM && defined(__CUDA_ARCH__)
#define YRT_WORLD_IN_SMEM 1
#else
#define YRT_WORLD_IN_SMEM 0
#endif
#define YRT_WORLD_WORDS 13    /* o, d, 1/d, -o/d, pad */

namespace yrt {

struct TraceCounters {   // optional per-ray work counters (roofline inputs), host_emu / debug kernels
    int box_tests, prim_tests, inst_entries, max_stack;
    int slab_false_rejects;   // boxes the reference's slab test accepts but the fused one rejects (must stay 0)
    int slab_extra_accepts;   // the other way round (harmless, costs a visit)
    int tlas_box_tests;       // part of box_tests spent in the instance tree
};

// test the prims of one BLAS leaf; returns true if any was hit (tmax/hit updated)
template <bool ANY>
YRT_HD bool leaf_prims(const SceneView& sv, int kind, int first, int count, const ray3& lray_in, float& tmax,
                       int si, HitRec& hit, TraceCounters* ctr) {
    bool any_hit = false;
    ray3 lray = lray_in;
    for (int k = first; k < first + count; k++) {
        lray.tmax = tmax;
        const float4* pr = sv.prim_recs + 3 * (size_t)k;
        float4 q0 = ld4(pr), q1 = ld4(pr + 1);
        float t, a = 0.f, b = 0.f;
        bool h;
        if (ctr) ctr->prim_tests++;
        if (kind == 0) {
            float4 q2 = ld4(pr + 2);
            h = intersect_triangle_edges(lray, xyz(q0), xyz(q1), xyz(q2), t, a, b);   // record = (v0, e1, e2)
        } else if (kind == 1) {
            float4 q2 = ld4(pr + 2);
            h = intersect_line(lray, xyz(q0), xyz(q1), q1.w, q2.x, t, a);
        } else {
            h = intersect_point(lray, xyz(q0), q1.x, t);
        }
        if (h && !ANY && hit.si >= 0 && t == tmax) {
            // exact-distance tie with the current closest hit: the reference keeps whichever candidate
            // its own traversal visits LAST (scene.cpp:256 accepts t == tmax and overwrites); reproduce
            // that with the precomputed visit ranks instead of depending on our visit order
            int ri_new = sv.inst_rank[si], ri_old = sv.inst_rank[hit.si];
            h = ri_new > ri_old || (ri_new == ri_old && sv.prim_rank[k] > sv.prim_rank[hit.prim]);
        }
        if (h) {
            tmax = t;
            hit.si = si;
            hit.prim = k;
            hit.w1 = a;
            hit.w2 = b;
            hit.dist = t;
            any_hit = true;
            if (ANY) return true;
        }
    }
    return any_hit;
}

YRT_HD vec3 inv3(const vec3& d) { return mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); }   // scene.cpp:372

// Reciprocal direction for the FUSED slab test only (never for Tracer<ANY, true>, which applies the reference's formula):
// that test is a conservative cull, so 1/d may carry the 1 ulp (2u) error of MUFU.RCP instead of the correctly rounded
// IEEE quotient (8 instructions per component) — the accept factor in slab_test_node budgets for it (yrt_math.cuh).
// Host build (tests/host_emu): the correctly rounded quotient pushed one full ulp up or down, direction picked from
// the operand's bits, so that the audit of every box test covers the worst case of the device's approximation.
YRT_HD float rcp_slab(float x) {
#if !YRT_APPROX_RCP
    return 1.0f / x;
#elif defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    float r = 1.0f / x;
    if (!(fabsf(r) <= 3.0e38f) || r == 0.0f) return r;
    int b = float_as_int(r);
    return int_as_float((float_as_int(x) * 0x9E3779B1u) & 0x10000u ? b + 1 : b - 1);
#endif
}
YRT_HD vec3 inv3_slab(const vec3& d) { return mk3(rcp_slab(d.x), rcp_slab(d.y), rcp_slab(d.z)); }

// Resumable traversal state of one ray.  Closest hit (ANY=false) keeps shrinking tmax; any hit
// (ANY=true) stops at the first accepted primitive (scene.cpp:414,425,436,473).
//
// Loop shape ("while-while"): nodes() walks internal nodes — TLAS and BLAS nodes share the same code
// and one index space, only the ray registers differ — until the lane holds a leaf; leaf() then runs
// the leaf code once (instance entry: transform the ray; element leaf: primitive tests).  With an
// if-if loop the node path and the leaf paths would be issued in every iteration for partial warps;
// the kernels are issue bound, so that matters (profiles/).  Being resumable lets the persistent
// kernels hand a finished lane a new ray while the rest of the warp keeps going.
template <bool ANY, bool EXACT = false>
struct Tracer {
#if YRT_WORLD_IN_SMEM
    // the world-space ray is needed again only when the lane leaves an instance: it waits in shared memory
    // (word w of thread t at [w * blockDim.x + t]: conflict-free), kernels launch with YRT_WORLD_WORDS * 4 bytes per thread
    __device__ __forceinline__ static float* wsave() {
        extern __shared__ float yrt_world_save[];
        return yrt_world_save + threadIdx.x;
    }
    __device__ __forceinline__ void save_world() {
        float* w = wsave();
        const unsigned T = blockDim.x;
        w[0] = o.x; w[T] = o.y; w[2 * T] = o.z; w[3 * T] = d.x; w[4 * T] = d.y; w[5 * T] = d.z;
        w[6 * T] = sr.invd.x; w[7 * T] = sr.invd.y; w[8 * T] = sr.invd.z;
        w[9 * T] = sr.noi.x; w[10 * T] = sr.noi.y; w[11 * T] = sr.noi.z; w[12 * T] = sr.pad;
    }
    __device__ __forceinline__ void restore_world() {
        const float* w = wsave();
        const unsigned T = blockDim.x;
        o = mk3(w[0], w[T], w[2 * T]); d = mk3(w[3 * T], w[4 * T], w[5 * T]);
        sr.invd = mk3(w[6 * T], w[7 * T], w[8 * T]);
        sr.noi = mk3(w[9 * T], w[10 * T], w[11 * T]);
        sr.ainv = mk3(fabsf(sr.invd.x), fabsf(sr.invd.y), fabsf(sr.invd.z));
        sr.pad = w[12 * T];
    }
    __device__ __forceinline__ const vec3& world_o() const { return o; }
    __device__ __forceinline__ const vec3& world_d() const { return d; }
#else
    vec3 wo, wd;          // world-space ray
    slabray wsr;
    YRT_HD void save_world() { wo = o; wd = d; wsr = sr; }
    // At the top level only the slab-test operands are read (the fused test needs neither o nor d, and the instance entry
    // transforms wo / wd): (o, d) may keep the last instance's values there — six moves less per instance exit.  The
    // reference's slab formula (EXACT) reads o.
    YRT_HD void restore_world() {
#if YRT_RESTORE_OD
        o = wo; d = wd;
#else
        if (EXACT) o = wo;
#endif
        sr = wsr;
    }
    YRT_HD const vec3& world_o() const { return wo; }
    YRT_HD const vec3& world_d() const { return wd; }
#endif
    vec3 o, d;            // ray in the current space (world, or local to instance `si`)
    slabray sr;
    float tmin, tmax;     // tmin is copied unchanged into every instance space (vmath.h:277)
    int cur, si, kind;
    int* sp;              // next free stack slot; stack[0] holds a YRT_REF_DONE guard, so a pop needs no emptiness test
#if YRT_STACK_TOP_REG
    int tos;              // top of the stack (the entries below it are in memory)
#endif
    bool top, found;
    HitRec hit;

    YRT_HD bool done() const { return cur == YRT_REF_DONE; }

    YRT_HD void begin(const SceneView& sv, const ray3& wray, int* stack) {
        hit.si = -1; hit.prim = -1; hit.w1 = hit.w2 = 0.f; hit.dist = 0.f;
        o = wray.o; d = wray.d;
        sr = make_slabray(o, EXACT ? inv3(d) : inv3_slab(d));
        save_world();
        tmin = wray.tmin; tmax = wray.tmax;
        stack[0] = YRT_REF_DONE; sp = stack + 1;
#if YRT_STACK_TOP_REG
        tos = YRT_REF_DONE;   // logical stack = [guard in memory, DONE in the register]: the last pop reads the guard, never below it
#endif
        si = -1; kind = 0; top = true; found = false;
        cur = sv.n_active_instances > 0 ? sv.tlas_root : YRT_REF_DONE;
    }

    // pop the next reference.  Leaving an instance (sentinel) restores the world-space ray and pops once more: only one
    // sentinel is ever on the stack (instances are entered from the top level only), and the guard below everything
    // (YRT_REF_DONE) ends the traversal without an emptiness test.
    YRT_HD int take() {
#if YRT_STACK_TOP_REG
        int r = tos;
        tos = *--sp;
        return r;
#else
        return *--sp;
#endif
    }
    YRT_HD void push(int ref) {
#if YRT_STACK_TOP_REG
        *sp++ = tos;
        tos = ref;
#else
        *sp++ = ref;
#endif
    }
    YRT_HD void pop() {
        cur = take();
        if (cur == YRT_REF_SENTINEL) {
            top = true;
            restore_world();
            cur = take();
        }
    }
    // a node whose visit is postponed: start pulling its record towards the SM
    YRT_HD void prefetch_node(const SceneView& sv, int ref) {
#if YRT_PREFETCH_PUSH && defined(__CUDA_ARCH__)
        if (ref >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(sv.nodes + YRT_NODE_STRIDE * (size_t)ref));
#else
        (void)sv; (void)ref;
#endif
    }

    // one internal node: test both child boxes against the current ray and current tmax, near child first
    YRT_HD void visit(const SceneView& sv, int* stack, TraceCounters* ctr) {
        float4 q0, q1, q2, q3;
        node_load(sv.nodes, cur, q0, q1, q2, q3);
        float e0, e1;
        bool h0, h1;
        if (EXACT) {
            // rays (nearly) parallel to an axis plane: |invd| is huge there and so is the per-box pad of the fused
            // test (it would accept half the scene); they take the reference's own formula on the stored box instead
            raysigns sgn = signs_of(sr.invd);
            nodebox b0 = node_child(q0, q1, q2, 0), b1 = node_child(q0, q1, q2, 1);
            h0 = intersect_check_bbox(o, sr.invd, sgn, tmin, tmax, b0.cx - b0.hx, b0.cy - b0.hy, b0.cz - b0.hz, b0.cx + b0.hx, b0.cy + b0.hy, b0.cz + b0.hz, e0);
            h1 = intersect_check_bbox(o, sr.invd, sgn, tmin, tmax, b1.cx - b1.hx, b1.cy - b1.hy, b1.cz - b1.hz, b1.cx + b1.hx, b1.cy + b1.hy, b1.cz + b1.hz, e1);
        } else {
            slab_test_node(sr, tmin, tmax, q0, q1, q2, h0, h1, e0, e1);
        }
        if (ctr) {   // host-side audit against the reference's own test
            ctr->box_tests += 2;
            if (top) ctr->tlas_box_tests += 2;
            float e;
            raysigns sgn = signs_of(sr.invd);   // the reference's test on the stored box [c-h, c+h] (a superset of the true box)
            nodebox b0 = node_child(q0, q1, q2, 0), b1 = node_child(q0, q1, q2, 1);
            const vec3 ao = top ? world_o() : o;   // (o is not restored at the top level)
            bool r0 = intersect_check_bbox(ao, sr.invd, sgn, tmin, tmax, b0.cx - b0.hx, b0.cy - b0.hy, b0.cz - b0.hz, b0.cx + b0.hx, b0.cy + b0.hy, b0.cz + b0.hz, e);
            bool r1 = intersect_check_bbox(ao, sr.invd, sgn, tmin, tmax, b1.cx - b1.hx, b1.cy - b1.hy, b1.cz - b1.hz, b1.cx + b1.hx, b1.cy + b1.hy, b1.cz + b1.hz, e);
            ctr->slab_false_rejects += (r0 && !h0) + (r1 && !h1);
            ctr->slab_extra_accepts += (!r0 && h0) + (!r1 && h1);
        }
        int c0, c1;
        node_refs(q1, q3, c0, c1);
#if YRT_VISIT_V2
        // one select for the next node, one predicated push when both children are entered, one branch for the pop
        if (h0 || h1) {
            int nxt = h0 ? c0 : c1;
            if (h0 && h1) {
                bool swap = (ANY && YRT_ANY_UNORDERED) ? false : (e1 < e0);   // near child first
                const int later = swap ? c0 : c1;
                push(later);
                prefetch_node(sv, later);
                if (ctr && (int)(sp - stack) > ctr->max_stack) ctr->max_stack = (int)(sp - stack);
                if (swap) nxt = c1;
            }
            cur = nxt;
        } else {
            pop();
        }
#else
        if (h0 && h1) {
            bool swap = (ANY && YRT_ANY_UNORDERED) ? false : (e1 < e0);   // near child first
            push(swap ? c0 : c1);
            if (ctr && (int)(sp - stack) > ctr->max_stack) ctr->max_stack = (int)(sp - stack);
            cur = swap ? c1 : c0;
        } else if (h0) {
            cur = c0;
        } else if (h1) {
            cur = c1;
        } else {
            pop();
        }
#endif
    }

    // internal nodes until the lane holds a leaf
    YRT_HD void nodes(const SceneView& sv, int* stack, TraceCounters* ctr) {
        while (cur >= 0) visit(sv, stack, ctr);
    }

    // one leaf (cur < 0 and not done)
    YRT_HD void leaf(const SceneView& sv, int* stack, TraceCounters* ctr) {
        int first = leaf_first(cur), count = leaf_count(cur);
        if (top) {
            // TLAS leaf: enter its first instance, keep the rest for later
            if (count > 1) push(make_leaf_ref(first + 1, count - 1));
            const float4* ir = sv.inst_recs + 4 * (size_t)first;
            float4 q0, q1, q2, q3;
            ld8(ir, q0, q1);
            ld8(ir + 2, q2, q3);
            frame3 f;
            f.x = xyz(q0); f.y = xyz(q1); f.z = xyz(q2); f.o = xyz(q3);
            vec3 lo = transform_point_inverse(f, world_o());   // transform_ray_inverse, scene.cpp:468
            d = transform_direction_inverse(f, world_d());
            o = lo;
            sr = make_slabray(o, EXACT ? inv3(d) : inv3_slab(d));
            si = first;
            kind = ((unsigned)float_as_int(q3.w)) >> 28;
            top = false;
            push(YRT_REF_SENTINEL);
            if (ctr) { ctr->inst_entries++; if ((int)(sp - stack) > ctr->max_stack) ctr->max_stack = (int)(sp - stack); }
            cur = float_as_int(q0.w);   // BLAS root ref of the instance's shape
        } else {
            ray3 lray;
            lray.o = o; lray.d = d; lray.tmin = tmin; lray.tmax = tmax;
            if (leaf_prims<ANY>(sv, kind, first, count, lray, tmax, si, hit, ctr)) {
                found = true;
                if (ANY) { cur = YRT_REF_DONE; return; }
            }
            pop();
        }
    }
};

// |invd| above which a ray uses the reference's slab formula (Tracer<ANY, true>): below it the fused test's per-box pad
// 8u max|invd| (hx+hy+hz) stays under 0.002 (hx+hy+hz)
#define YRT_EXACT_SLAB_INVD 4096.0f

template <bool ANY, bool EXACT>
YRT_HD bool trace_ray_impl(const SceneView& sv, const ray3& wray, HitRec& hit, int* stack, TraceCounters* ctr) {
    Tracer<ANY, EXACT> t;
    t.begin(sv, wray, stack);
    for (;;) {
        t.nodes(sv, stack, ctr);
        if (t.done()) break;
        t.leaf(sv, stack, ctr);
    }
    hit = t.hit;
    return t.found;
}

template <bool ANY>
YRT_HD bool trace_ray(const SceneView& sv, const ray3& wray, HitRec& hit, int* stack, TraceCounters* ctr) {
    // decided on the world-space direction (instance frames of the configs are pure translations, so the local
    // direction is the same; a rotated instance may still meet a large pad — slower, never wrong).
    // |1/d| > 4096  <=>  |d| < 2^-12 (the correctly rounded quotient is monotonic and exact at the power of two)
    float ax = fabsf(wray.d.x), ay = fabsf(wray.d.y), az = fabsf(wray.d.z);
    float m = fminf(fminf(ax, ay), az), big = fmaxf(fmaxf(ax, ay), az);
    // (directions beyond 1e30 would underflow the flush-to-zero MUFU reciprocal: exact path as well; NaN goes there too)
    if (!(m >= 1.0f / YRT_EXACT_SLAB_INVD && big <= 1.0e30f)) return trace_ray_impl<ANY, true>(sv, wray, hit, stack, ctr);
    return trace_ray_impl<ANY, false>(sv, wray, hit, stack, ctr);
}

}  // namespace yrt
