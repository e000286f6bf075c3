// yrt_packet.cuh — warp-cooperative ("packet") traversal: the 32 rays of a warp walk ONE tree path.
//
// The rays a warp holds are neighbouring samples (16 samples of 2 pixels; their shadow rays towards one light),
// so they visit almost the same nodes.  With one traversal per lane (Tracer, yrt_trace.cuh) the warp still pays
// for its longest lane in every phase: ncu shows 20-22 of 32 lanes active in the node loop (profiles/r1d).  Here
// the traversal state — current node, stack, level, instance — is WARP-UNIFORM: a node is visited if ANY lane's
// ray enters its box (warp vote), every lane tests its own ray against both child boxes and against the
// elements of every leaf the warp reaches, the stack lives once per warp in shared memory, node / element /
// instance records are fetched with one broadcast load per warp instead of 32 gathers, and there is no
// divergence to reconverge from.  Per ray the set of elements tested becomes a SUPERSET of what its own
// traversal would test; the element tests, the instance transform, the closest-hit rule and the tie rule are the
// same exact per-lane code, so results are unchanged (GPU parity tests).
// Device only (warp votes); incoherent warps are better served by Tracer — see use_packet() in yrt_render.cu.
#pragma once
#include "yrt_trace.cuh"

namespace yrt {

#if defined(__CUDACC__)

#define YRT_WSTACK YRT_STACK_CAP   /* per-warp stack entries in shared memory (the build checks depth + 4 <= YRT_STACK_CAP) */

template <bool ANY>
__device__ __forceinline__ void trace_packet(const SceneView& sv, const ray3& wray, bool alive, HitRec& hit, bool& found,
                                             int* __restrict__ wstack, int lane) {
    const unsigned FULL = 0xffffffffu;
    hit.si = -1; hit.prim = -1; hit.w1 = hit.w2 = 0.f; hit.dist = 0.f;
    found = false;
    if (sv.n_active_instances <= 0) return;

    // per-lane ray state
    const vec3 wo = wray.o, wd = wray.d;
    const slabray wsr = make_slabray(wo, inv3_slab(wd));
    vec3 o = wo, d = wd;
    slabray sr = wsr;
    const float tmin = wray.tmin;
    float tmax = alive ? wray.tmax : -FLT_MAX;     // a dead lane fails every slab test and every element test
    // warp-uniform traversal state
    int sp = 0, cur = sv.tlas_root, si = -1, kind = 0;
    bool top = true;

// warp-uniform pop; leaving an instance (sentinel) restores every lane's world-space ray
#define YRT_WPOP()                                   \
    __syncwarp();                                    \
    for (;;) {                                       \
        if (sp == 0) { cur = YRT_REF_DONE; break; }  \
        cur = wstack[--sp];                          \
        if (cur != YRT_REF_SENTINEL) break;          \
        top = true;                                  \
        o = wo; d = wd; sr = wsr;                    \
    }

    for (;;) {
        // ---- internal nodes ----
        while (cur >= 0) {
            float4 q0, q1, q2, q3;
            node_load(sv.nodes, cur, q0, q1, q2, q3);   // same address in every lane: one broadcast each
            float e0, e1;
            bool h0, h1;
            slab_test_node(sr, tmin, tmax, q0, q1, q2, h0, h1, e0, e1);
            unsigned m0 = __ballot_sync(FULL, h0), m1 = __ballot_sync(FULL, h1);
            int c0, c1;
            node_refs(q1, q3, c0, c1);
            if (m0 && m1) {
                bool swap = false;
                if (!ANY) {   // near child first, by majority of the lanes that enter both (else of all entering lanes)
                    unsigned p1 = __ballot_sync(FULL, h0 && h1 && e1 < e0);
                    unsigned both = m0 & m1;
                    swap = both ? (2 * __popc(p1) > __popc(both)) : (__popc(m1) > __popc(m0));
                }
                if (lane == 0) wstack[sp] = swap ? c0 : c1;
                sp++;
                cur = swap ? c1 : c0;
            } else if (m0) {
                cur = c0;
            } else if (m1) {
                cur = c1;
            } else {
                YRT_WPOP();
            }
        }
        if (cur == YRT_REF_DONE) break;
        // ---- leaf ----
        int first = leaf_first(cur), count = leaf_count(cur);
        if (top) {
            // TLAS leaf: the warp enters its first instance (every lane transforms its own ray), the rest waits on the stack
            if (count > 1) { if (lane == 0) wstack[sp] = make_leaf_ref(first + 1, count - 1); sp++; }
            const float4* ir = sv.inst_recs + 4 * (size_t)first;
            float4 q0 = ld4(ir), q1 = ld4(ir + 1), q2 = ld4(ir + 2), q3 = ld4(ir + 3);
            frame3 f;
            f.x = xyz(q0); f.y = xyz(q1); f.z = xyz(q2); f.o = xyz(q3);
            o = transform_point_inverse(f, wo);          // transform_ray_inverse, scene.cpp:468
            d = transform_direction_inverse(f, wd);
            sr = make_slabray(o, inv3_slab(d));
            si = first;
            kind = ((unsigned)float_as_int(q3.w)) >> 28;
            top = false;
            if (lane == 0) wstack[sp] = YRT_REF_SENTINEL;
            sp++;
            cur = float_as_int(q0.w);                    // BLAS root ref of the instance's shape
        } else {
            ray3 lray;
            lray.o = o; lray.d = d; lray.tmin = tmin; lray.tmax = tmax;
            if (leaf_prims<ANY>(sv, kind, first, count, lray, tmax, si, hit, nullptr)) {
                found = true;
                if (ANY) tmax = -FLT_MAX;                // this lane is done; it no longer steers the packet
            }
            if (ANY && __ballot_sync(FULL, tmax > -FLT_MAX) == 0u) break;   // every live lane has its answer
            YRT_WPOP();
        }
    }
#undef YRT_WPOP
}

#endif  // __CUDACC__

}  // namespace yrt
