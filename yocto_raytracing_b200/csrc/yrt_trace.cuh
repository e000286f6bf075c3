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
#include "yrt_pgrid.cuh"
#include "yrt_scene.cuh"

#ifndef YRT_POP_CULL
#define YRT_POP_CULL 1   /* closest hit: a stack entry carries the entry distance of its box and is dropped at pop time if the hit found since then is nearer */
#endif
#ifndef YRT_ANY_UNORDERED
#define YRT_ANY_UNORDERED 1  /* any-hit rays: children are entered in stored order, no entry-distance compare (the answer is order independent) */
#endif

namespace yrt {

struct TraceCounters {   // optional per-ray work counters (roofline inputs): host emulation and the -DYRT_COUNTERS build of the library
    int box_tests, prim_tests, inst_entries, max_stack;
    int slab_false_rejects;   // boxes the reference's slab test accepts but the fused one rejects (must stay 0; host audit only)
    int slab_extra_accepts;   // the other way round (harmless, costs a visit; host audit only)
    int tlas_box_tests;       // part of box_tests spent in the instance tree
    int node_visits;          // node records fetched (2 or up to 4 box tests each)
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
//
// GRID: the ray belongs to an apex grid in list form (yrt_pgrid.cuh, the camera's) — its instance level is the candidate
// list of its cell: entries are taken in list order (nearest first), each one's box is tested with the same slab test,
// and the instances it touches are entered; only the shapes' trees are walked.  The top level is the state
// cur == YRT_REF_SENTINEL ("take the next candidate"), reached at the start and whenever the walk of an instance ends.
template <bool ANY, bool EXACT = false, bool GRID = false>
struct Tracer {
    vec3 wo, wd;          // world-space ray
    slabray wsr;
    vec3 o, d;            // ray in the current space (world, or local to instance `si`)
    slabray sr;
    float tmin, tmax;     // tmin is copied unchanged into every instance space (vmath.h:277)
    int cur, si, kind;
    int* stk;             // the lane's stack (local memory)
    int sp;               // next free word of it; the first entry is a YRT_REF_DONE guard, so a pop needs no emptiness test
    bool top, found;
    int gi, gend;         // GRID: next and end entry of the cell's candidate list
    HitRec hit;

    YRT_HD bool done() const { return cur == YRT_REF_DONE; }

    // root: the node the walk starts from — the instance tree's root, or the root of the ray's cell in an apex grid
    // (yrt_pgrid.cuh: a chain of instance-level nodes over the few instances a ray of that cell can touch)
    YRT_HD void begin(const SceneView& sv, const ray3& wray, int* stack, int root) {
        hit.si = -1; hit.prim = -1; hit.w1 = hit.w2 = 0.f; hit.dist = 0.f;
        o = wray.o; d = wray.d;
        sr = make_slabray(o, EXACT ? inv3(d) : inv3_slab(d));
        wo = o; wd = d; wsr = sr;
        tmin = wray.tmin; tmax = wray.tmax;
        stk = stack; sp = 0;
        push(YRT_REF_DONE, -FLT_MAX);
        si = -1; kind = 0; top = true; found = false;
        gi = gend = 0;
        cur = sv.n_active_instances > 0 ? root : YRT_REF_DONE;
    }
    // GRID: the ray's candidates are entries [first, first + count) of its grid
    YRT_HD void begin_list(const SceneView& sv, const ray3& wray, int* stack, int first, int count) {
        begin(sv, wray, stack, YRT_REF_SENTINEL);
        gi = first; gend = first + count;
    }

    // Stack entries.  Closest hit (YRT_POP_CULL): (reference, entry distance of its box); the pop re-applies the accept
    // rule of the slab test with the tmax of NOW — exactly the decision the box test would take if it ran at pop time
    // (its exit-side operands have not changed), so a postponed subtree or instance that a nearer hit has made
    // irrelevant costs one compare instead of a visit.  Wide nodes need this more than binary ones: up to three siblings
    // wait on the stack per visit, and instance leaves are entered straight from it.  Any hit: the reference only.
    static constexpr int W = ANY ? YRT_WIDE_ANY : YRT_WIDE_CLOSEST;   // arity of the node records this ray kind walks
    static constexpr bool CULL = !ANY && YRT_POP_CULL;
    static constexpr int ENTRY = CULL ? 2 : 1;
    YRT_HD float accept_limit() const { return EXACT ? tmax * 1.00000024f : fmaf(tmax, YRT_SLAB_ACCEPT, sr.pad); }

    // pop the next reference.  Leaving an instance (sentinel) restores the world-space slab operands and pops on: only
    // one sentinel is ever on the stack (instances are entered from the top level only), and the guard below everything
    // (YRT_REF_DONE, entry distance -inf like the sentinel's) ends the traversal without an emptiness test.  At the top
    // level only the slab-test operands are read (the fused test needs neither o nor d, and the instance entry
    // transforms wo / wd), so (o, d) keep the last instance's values there; the reference's slab formula (EXACT) reads o.
    YRT_HD void pop_entry() {
        if (CULL) {
            // only the stack pointer moves inside the loop (no other state is live across its back edge)
            const float lim = accept_limit();
            do { sp -= 2; } while (!(int_as_float(stk[sp + 1]) <= lim));
            cur = stk[sp];
        } else {
            cur = stk[--sp];
        }
    }
    YRT_HD void pop() {
        pop_entry();
        if (cur == YRT_REF_SENTINEL) {
            top = true;
            if (EXACT) o = wo;
            sr = wsr;
            if (!GRID) pop_entry();   // GRID: cur stays the sentinel = "next candidate of the list" (next_candidate())
        }
    }
    YRT_HD void push(int ref, float e) {
        stk[sp] = ref;
        if (CULL) stk[sp + 1] = float_as_int(e);
        sp += ENTRY;
    }
    YRT_HD int depth_of(const int*) const { return sp / ENTRY; }

    // both boxes of one pair against the current ray and the current tmax
    YRT_HD void test_pair(const float4& p0, const float4& p1, const float4& p2, bool& h0, bool& h1, float& e0, float& e1, TraceCounters* ctr) {
        if (EXACT) {
            // rays (nearly) parallel to an axis plane: |invd| is huge there and so is the per-box pad of the fused
            // test (it would accept half the scene); they take the reference's own formula on the stored box instead
            raysigns sgn = signs_of(sr.invd);
            nodebox b0 = pair_child(p0, p1, p2, 0), b1 = pair_child(p0, p1, p2, 1);
            h0 = intersect_check_bbox(o, sr.invd, sgn, tmin, tmax, b0.cx - b0.hx, b0.cy - b0.hy, b0.cz - b0.hz, b0.cx + b0.hx, b0.cy + b0.hy, b0.cz + b0.hz, e0);
            h1 = intersect_check_bbox(o, sr.invd, sgn, tmin, tmax, b1.cx - b1.hx, b1.cy - b1.hy, b1.cz - b1.hz, b1.cx + b1.hx, b1.cy + b1.hy, b1.cz + b1.hz, e1);
        } else {
            slab_test_pair(sr, tmin, tmax, p0, p1, p2, h0, h1, e0, e1);
        }
#if !defined(__CUDA_ARCH__)
        if (ctr) {   // host-side audit against the reference's own test on the stored box [c-h, c+h] (a superset of the true box)
            float e;
            raysigns sgn = signs_of(sr.invd);
            nodebox b0 = pair_child(p0, p1, p2, 0), b1 = pair_child(p0, p1, p2, 1);
            const vec3 ao = top ? wo : o;   // (o is not restored at the top level)
            bool r0 = b0.hx >= 0.f && intersect_check_bbox(ao, sr.invd, sgn, tmin, tmax, b0.cx - b0.hx, b0.cy - b0.hy, b0.cz - b0.hz, b0.cx + b0.hx, b0.cy + b0.hy, b0.cz + b0.hz, e);
            bool r1 = b1.hx >= 0.f && intersect_check_bbox(ao, sr.invd, sgn, tmin, tmax, b1.cx - b1.hx, b1.cy - b1.hy, b1.cz - b1.hz, b1.cx + b1.hx, b1.cy + b1.hy, b1.cz + b1.hz, e);
            ctr->slab_false_rejects += (r0 && !h0) + (r1 && !h1);
            ctr->slab_extra_accepts += (!r0 && h0) + (!r1 && h1);
        }
#else
        (void)ctr;
#endif
    }

    // One wide node: the four child boxes against the current ray and current tmax.
    // Any hit: the entered children are taken in stored order.  Closest hit: nearest child first; the others are pushed
    // so that they come off the stack in the order [far child of the nearest child's pair | near child of the other
    // pair | far child of the other pair], the last two swapped if the entry distances say so — i.e. sorted by entry
    // distance except that a pair is never split by more than one position (a tournament, not a full sort).
    YRT_HD void visit4(const SceneView& sv, int* stack, TraceCounters* ctr) {
        const float4* n = sv.nodes4 + YRT_NODE_STRIDE(4) * (size_t)cur;
        float4 a0 = ld4(n), a1 = ld4(n + 1), a2 = ld4(n + 2), b0 = ld4(n + 3), b1 = ld4(n + 4), b2 = ld4(n + 5);
        int c0, c1, c2, c3;
        {
            float4 r = ld4(n + 6);
            c0 = float_as_int(r.x); c1 = float_as_int(r.y); c2 = float_as_int(r.z); c3 = float_as_int(r.w);
        }
        bool h0, h1, h2, h3;
        float e0, e1, e2, e3;
        test_pair(a0, a1, a2, h0, h1, e0, e1, ctr);
        test_pair(b0, b1, b2, h2, h3, e2, e3, ctr);
        if (ctr) {
            ctr->node_visits++;
            int nb = float_as_int(ld4(n + 7).x);
            ctr->box_tests += nb;
            if (top) ctr->tlas_box_tests += nb;
        }
        if (ANY && YRT_ANY_UNORDERED) {
            // entered children in stored order: the first becomes the next node, the others wait on the stack.
            // Branch-free: every entered child stores the candidate it displaces at *sp, and sp advances only if there
            // was one (a store that is not followed by an advance is overwritten or never read).
            int nxt = YRT_REF_DONE, adv = 0;
            if (h3) { nxt = c3; adv = 1; }
            if (h2) { stk[sp] = nxt; sp += adv; nxt = c2; adv = 1; }
            if (h1) { stk[sp] = nxt; sp += adv; nxt = c1; adv = 1; }
            if (h0) { stk[sp] = nxt; sp += adv; nxt = c0; adv = 1; }
            if (ctr && depth_of(stack) > ctr->max_stack) ctr->max_stack = depth_of(stack);
            if (adv) cur = nxt; else pop();
        } else {
            const float inf = int_as_float(0x7f800000);
            if (!h0) e0 = inf;
            if (!h1) e1 = inf;
            if (!h2) e2 = inf;
            if (!h3) e3 = inf;
            // tournament: winner and loser of each pair, then of the two winners
            bool sa = e1 < e0, sb = e3 < e2;
            int wa = sa ? c1 : c0, la = sa ? c0 : c1, wb = sb ? c3 : c2, lb = sb ? c2 : c3;
            float ewa = fminf(e0, e1), ela = fmaxf(e0, e1), ewb = fminf(e2, e3), elb = fmaxf(e2, e3);
            bool sf = ewb < ewa;
            int w = sf ? wb : wa;            // nearest entered child
            int m = sf ? wa : wb;            // winner of the other pair
            float em = fmaxf(ewa, ewb);
            int l1 = sf ? lb : la, l2 = sf ? la : lb;      // loser of the winner's pair, loser of the other pair
            float el1 = sf ? elb : ela, el2 = sf ? ela : elb;
            if (fminf(ewa, ewb) < inf) {
                // farthest first onto the stack: l2 | (m, l1) in entry order
                if (el2 < inf) push(l2, el2);
                bool swap = el1 < em;      // l1 nearer than m: m goes below it
                int x = swap ? m : l1, y = swap ? l1 : m;
                float ex = swap ? em : el1, ey = swap ? el1 : em;
                if (ex < inf) push(x, ex);
                if (ey < inf) push(y, ey);
                if (ctr && depth_of(stack) > ctr->max_stack) ctr->max_stack = depth_of(stack);
                cur = w;
            } else {
                pop();
            }
        }
    }

    // one binary node: test both child boxes against the current ray and current tmax, near child first
    YRT_HD void visit2(const SceneView& sv, int* stack, TraceCounters* ctr) {
        const float4* n = sv.nodes2 + YRT_NODE_STRIDE(2) * (size_t)cur;
        float4 q0 = ld4(n), q1 = ld4(n + 1), q2 = ld4(n + 2);
        int c0, c1;
        {
            float4 r = ld4(n + 3);
            c0 = float_as_int(r.x); c1 = float_as_int(r.y);
        }
        float e0, e1;
        bool h0, h1;
        test_pair(q0, q1, q2, h0, h1, e0, e1, ctr);
        if (ctr) {
            ctr->node_visits++;
            ctr->box_tests += 2;
            if (top) ctr->tlas_box_tests += 2;
        }
        // one select for the next node, one predicated push when both children are entered, one branch for the pop
        if (h0 || h1) {
            int nxt = h0 ? c0 : c1;
            if (h0 && h1) {
                bool swap = (ANY && YRT_ANY_UNORDERED) ? false : (e1 < e0);   // near child first
                push(swap ? c0 : c1, swap ? e0 : e1);
                if (ctr && depth_of(stack) > ctr->max_stack) ctr->max_stack = depth_of(stack);
                if (swap) nxt = c1;
            }
            cur = nxt;
        } else {
            pop();
        }
    }
    YRT_HD void visit(const SceneView& sv, int* stack, TraceCounters* ctr) {
        if (W == 4) visit4(sv, stack, ctr); else visit2(sv, stack, ctr);
    }

    // internal nodes until the lane holds a leaf
    YRT_HD void nodes(const SceneView& sv, int* stack, TraceCounters* ctr) {
        while (cur >= 0) visit(sv, stack, ctr);
    }

    // GRID, top level: candidates in list order until one's box is touched (-> enter it) or the list ends / the rest lies
    // beyond the ray's reach (-> done).  The .w of an entry's second quad is a lower bound of the distance from the apex to
    // anything of the instance a ray of this cell can reach; camera rays measure t from the apex: a candidate whose bound
    // exceeds tmax cannot be hit any more, and neither can the ones sorted behind it.
    YRT_HD void next_candidate(const SceneView& sv, const GridRef& g, int* stack, TraceCounters* ctr) {
        for (;;) {
            if (gi >= gend) { cur = YRT_REF_DONE; return; }
            const float4* e = g.entries + 2 * (size_t)gi;
            const float4 e0 = ld4(e), e1 = ld4(e + 1);
            gi++;
            if (e1.w > tmax) { cur = YRT_REF_DONE; return; }
            float te;
            if (ctr) { ctr->box_tests++; ctr->tlas_box_tests++; }
            if (slab_test_ch(wsr, tmin, tmax, e0.x, e0.y, e0.z, e1.x, e1.y, e1.z, te)) {
                enter_instance(sv, float_as_int(e0.w), stack, ctr);
                return;
            }
        }
    }

    // one leaf (cur < 0 and not done)
    YRT_HD void leaf(const SceneView& sv, int* stack, TraceCounters* ctr) {
        int first = leaf_first(cur), count = leaf_count(cur);
        if (top) {
            // TLAS leaf: enter its first instance, keep the rest for later
            if (count > 1) push(make_leaf_ref(first + 1, count - 1), -FLT_MAX);
            enter_instance(sv, first, stack, ctr);
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

    // the ray enters instance slot `first`: exact transform_ray_inverse (scene.cpp:468), then the root of the shape's tree
    YRT_HD void enter_instance(const SceneView& sv, int first, int* stack, TraceCounters* ctr) {
        {
            const float4* ir = sv.inst_recs + 4 * (size_t)first;
            float4 q0, q1, q2, q3;
            q0 = ld4(ir); q1 = ld4(ir + 1); q2 = ld4(ir + 2); q3 = ld4(ir + 3);
            frame3 f;
            f.x = xyz(q0); f.y = xyz(q1); f.z = xyz(q2); f.o = xyz(q3);
            vec3 lo = transform_point_inverse(f, wo);   // transform_ray_inverse, scene.cpp:468
            d = transform_direction_inverse(f, wd);
            o = lo;
            sr = make_slabray(o, EXACT ? inv3(d) : inv3_slab(d));
            si = first;
            kind = ((unsigned)float_as_int(q3.w)) >> 28;
            top = false;
            push(YRT_REF_SENTINEL, -FLT_MAX);
            if (ctr) { ctr->inst_entries++; if (depth_of(stack) > ctr->max_stack) ctr->max_stack = depth_of(stack); }
            cur = float_as_int(q0.w);   // BLAS root ref of the instance's shape
        }
    }
};

// |invd| above which a ray uses the reference's slab formula (Tracer<ANY, true>): below it the fused test's per-box pad
// 8u max|invd| (hx+hy+hz) stays under 0.002 (hx+hy+hz)
#define YRT_EXACT_SLAB_INVD 4096.0f

template <bool ANY, bool EXACT>
YRT_HD bool trace_ray_impl(const SceneView& sv, const ray3& wray, HitRec& hit, int* stack, TraceCounters* ctr, int root) {
    Tracer<ANY, EXACT> t;
    t.begin(sv, wray, stack, root);
    for (;;) {
        t.nodes(sv, stack, ctr);
        if (t.done()) break;
        t.leaf(sv, stack, ctr);
    }
    hit = t.hit;
    return t.found;
}

// walk from `root` (see Tracer::begin)
template <bool ANY>
YRT_HD bool trace_ray_from(const SceneView& sv, const ray3& wray, HitRec& hit, int* stack, TraceCounters* ctr, int root) {
    // decided on the world-space direction (instance frames of the configs are pure translations, so the local
    // direction is the same; a rotated instance may still meet a large pad — slower, never wrong).
    // |1/d| > 4096  <=>  |d| < 2^-12 (the correctly rounded quotient is monotonic and exact at the power of two)
    float ax = fabsf(wray.d.x), ay = fabsf(wray.d.y), az = fabsf(wray.d.z);
    float m = fminf(fminf(ax, ay), az), big = fmaxf(fmaxf(ax, ay), az);
    // (directions beyond 1e30 would underflow the flush-to-zero MUFU reciprocal: exact path as well; NaN goes there too)
    if (!(m >= 1.0f / YRT_EXACT_SLAB_INVD && big <= 1.0e30f)) return trace_ray_impl<ANY, true>(sv, wray, hit, stack, ctr, root);
    return trace_ray_impl<ANY, false>(sv, wray, hit, stack, ctr, root);
}
template <bool ANY>
YRT_HD bool trace_ray(const SceneView& sv, const ray3& wray, HitRec& hit, int* stack, TraceCounters* ctr) {
    return trace_ray_from<ANY>(sv, wray, hit, stack, ctr, sv.tlas_root);
}

// rays the fused slab test cannot serve (see trace_ray_from): they walk from the tree's root with the reference's formula
YRT_HD bool ray_needs_exact_slabs(const ray3& wray) {
    float ax = fabsf(wray.d.x), ay = fabsf(wray.d.y), az = fabsf(wray.d.z);
    float m = fminf(fminf(ax, ay), az), big = fmaxf(fmaxf(ax, ay), az);
    return !(m >= 1.0f / YRT_EXACT_SLAB_INVD && big <= 1.0e30f);
}

// Scenes with non-rigid instance frames (RefTlas, yrt_scene.cuh): intersect_bvh(scene…) of src/scene.cpp:446-479 restated on a
// copy of the reference's own instance tree — node popped, slab test of scene.cpp:371-383 on the node's exact box with the
// tmax of now, inner nodes push (start, start + 1), leaves run their instances in array order; each instance is entered
// with transform_ray_inverse and the current tmax, and a hit's LOCAL distance becomes the new tmax (scene.cpp:468-470).
// The shapes' trees are the LBVH's: inside one instance the closest element with t <= tmax (ties by the reference's visit
// rank) does not depend on the tree.  `tstack`: YRT_REF_TLAS_STACK ints (the reference's own node_stack has 64).
#define YRT_REF_TLAS_STACK 64
YRT_HD int ldi(const int* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
template <bool ANY, bool EXACT>
YRT_HD bool trace_ray_ref_impl(const SceneView& sv, const RefTlas& rt, const ray3& wray, HitRec& hit, int* stack, int* tstack, TraceCounters* ctr) {
    Tracer<ANY, EXACT> t;
    t.begin(sv, wray, stack, YRT_REF_DONE);
    const vec3 invd = inv3(wray.d);                     // scene.cpp:372 (the same three quotients on every call)
    const raysigns sgn = signs_of(invd);
    int tsp = 0;
    if (rt.n_nodes > 0) tstack[tsp++] = 0;
    while (tsp) {
        const float4* n = rt.nodes + 2 * (size_t)tstack[--tsp];
        const float4 q0 = ld4(n), q1 = ld4(n + 1);
        float te;
        if (ctr) { ctr->box_tests++; ctr->tlas_box_tests++; ctr->node_visits++; }
        if (!intersect_check_bbox(wray.o, invd, sgn, t.tmin, t.tmax, q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, te)) continue;
        const int a = float_as_int(q0.w), b = float_as_int(q1.w);
        if (a >= 0) {
            if (tsp + 2 > YRT_REF_TLAS_STACK) continue;   // (deeper than the reference's own stack: it would have overrun it)
            tstack[tsp++] = a;
            tstack[tsp++] = b;
            continue;
        }
        for (int i = ~a; i < ~a + b; i++) {
            const int slot = ldi(rt.slot_of_inst + ldi(rt.leaf_inst + i));
            if (slot < 0) continue;                     // a shape without elements: its tree is one empty leaf, never a hit
            t.sp = 0;
            t.push(YRT_REF_DONE, -FLT_MAX);
            t.enter_instance(sv, slot, stack, ctr);
            for (;;) {
                t.nodes(sv, stack, ctr);
                if (t.done()) break;
                t.leaf(sv, stack, ctr);
            }
            if (ANY && t.found) { hit = t.hit; return true; }
        }
    }
    hit = t.hit;
    return t.found;
}
template <bool ANY>
YRT_HD bool trace_ray_ref(const SceneView& sv, const RefTlas& rt, const ray3& wray, HitRec& hit, int* stack, int* tstack, TraceCounters* ctr) {
    if (ray_needs_exact_slabs(wray)) return trace_ray_ref_impl<ANY, true>(sv, rt, wray, hit, stack, tstack, ctr);
    return trace_ray_ref_impl<ANY, false>(sv, rt, wray, hit, stack, tstack, ctr);
}

YRT_HD int ld_root(const int* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// a ray of an apex grid in list form whose cell holds the candidates [first, first + count)
template <bool ANY>
YRT_HD bool trace_ray_list(const SceneView& sv, const ray3& wray, const GridRef& g, int first, int count, HitRec& hit, int* stack, TraceCounters* ctr) {
    Tracer<ANY, false, true> t;
    t.begin_list(sv, wray, stack, first, count);
    for (;;) {
        t.nodes(sv, stack, ctr);
        if (t.done()) break;
        if (t.cur == YRT_REF_SENTINEL) t.next_candidate(sv, g, stack, ctr);
        else t.leaf(sv, stack, ctr);
    }
    hit = t.hit;
    return t.found;
}

YRT_HD int2 ld_cell(const int2* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// closest hit of the camera ray of pixel (i, j): the candidate list of the pixel's cell replaces the instance tree when
// the frame has a camera grid and the cell is served
YRT_HD void trace_camera_ray(const SceneView& sv, const GridRef& cg, const ray3& ray, int i, int j, HitRec& h, int* stack, TraceCounters* ctr) {
    if (cg.nx > 0) {
        const int2 c = ld_cell(cg.cells + (size_t)(j >> cg.shift) * cg.nx + (i >> cg.shift));
        if (c.y >= 0 && !ray_needs_exact_slabs(ray)) {
            trace_ray_list<false>(sv, ray, cg, c.x, c.y, h, stack, ctr);
            return;
        }
    }
    trace_ray<false>(sv, ray, h, stack, ctr);
}

// occlusion of the shadow ray `sr` towards light k: starts at the root of the cell that -d falls into, seen from the light
YRT_HD bool trace_shadow_ray(const SceneView& sv, const LightGrids& lg, int k, const ray3& sr, HitRec& hr, int* stack, TraceCounters* ctr) {
    int root = sv.tlas_root;
    if (k < YRT_MAX_LIGHT_GRIDS && lg.g[k].nx > 0) root = ld_root(lg.g[k].roots + pgrid_cube_cell(-sr.d, lg.g[k].nx));
    return trace_ray_from<true>(sv, sr, hr, stack, ctr, root);
}

}  // namespace yrt
