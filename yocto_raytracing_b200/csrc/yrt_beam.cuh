// yrt_beam.cuh — the instance tree walked ONCE PER WARP: a conservative beam around the warp's 32 rays against a 32-wide
// hierarchy over the instances, with the lanes spread over the CHILDREN of a node instead of over the rays.
//
// Why: the 32 rays of a warp are neighbours (the 16 samples of two pixels; their shadow rays towards one light), so they
// reach the same handful of instances — 3.4 per warp of camera rays, 2.0 per warp of shadow rays in the headline scene
// (tools/beam_stats.py) — yet every lane walked the whole instance tree on its own: 18 binary node visits per camera ray,
// 6 four-wide visits per shadow ray, more than half of all box tests (profiles/r2_*).  Here the warp computes the bounds
// of its rays' origins and reciprocal directions (the beam), descends the 32-wide levels with one lane per child box
// (about 5 node visits per warp, lanes fully used), and ends with the list of instances the beam touches; each lane then
// tests only those few instance boxes against its OWN ray (the same fused slab test as before, same accept rule) and
// pushes the ones it enters; from there on everything is the per-lane traversal of yrt_trace.cuh.
//
// Parity: the beam may only ADD candidates.  A lane enters an instance iff its own slab test on that instance's box
// accepts — the test it would have applied at the leaf of the instance tree — so the set of instances a ray enters is the
// set of instance boxes its own test accepts, a superset of what the reference enters (SURVEY finding 4: visit order and
// tree shape are free; src/scene.cpp:446-479).  The beam test is an interval version of that slab test with slack far
// beyond the fused test's error bounds, on boxes that contain the lanes' (inflated) boxes.
// Warps whose rays do not form a beam (direction signs differ, a component near zero, too many candidates) fall back to
// the per-lane walk of the instance tree.
//
// Scalar pieces are __host__ __device__ and shared with tests/host_emu, which runs the same algorithm with loops in place
// of the warp intrinsics.
#pragma once
#include <algorithm>

#include "yrt_scene.cuh"

namespace yrt {

#define YRT_BEAM_FANOUT 32
#define YRT_BEAM_MAX_LEVELS 5          /* 32^5 instances */
#define YRT_BEAM_CAND_CAP 32           /* candidates per warp; more = fall back to the per-lane instance tree */
#define YRT_BEAM_STACK 160             /* warp-level stack of (level, group) entries: <= 31 per level + 32 */

// Bounds of a warp's rays, per axis mirrored so that every reciprocal direction is positive.
struct Beam {
    float sx, sy, sz;        // +1 / -1: the mirror applied to coordinates of that axis
    vec3 omin, omax;         // bounds of the mirrored origins
    vec3 imin, imax;         // bounds of the mirrored reciprocal directions (all > 0)
    float tmin, tmax;        // min of the rays' tmin, max of their tmax
    float pad;               // absolute slack: 4 x the largest per-ray pad of the fused slab test
};

// what one ray contributes to the beam (its mirrored origin / reciprocal direction); valid = this ray may be part of a beam
struct BeamLane { vec3 o, i; float tmin, tmax, pad; };
YRT_HD bool beam_lane_terms(const ray3& r, const slabray& sr, bool neg_x, bool neg_y, bool neg_z, BeamLane& b) {
    // same criterion as trace_ray's choice of the reference's slab formula: such rays never join a beam
    float ax = fabsf(r.d.x), ay = fabsf(r.d.y), az = fabsf(r.d.z);
    float m = fminf(fminf(ax, ay), az), big = fmaxf(fmaxf(ax, ay), az);
    if (!(m >= 1.0f / 4096.0f && big <= 1.0e30f)) return false;
    if ((sr.invd.x < 0.f) != neg_x || (sr.invd.y < 0.f) != neg_y || (sr.invd.z < 0.f) != neg_z) return false;
    b.o = mk3(neg_x ? -r.o.x : r.o.x, neg_y ? -r.o.y : r.o.y, neg_z ? -r.o.z : r.o.z);
    b.i = sr.ainv;           // |1/d| = the mirrored reciprocal
    b.tmin = r.tmin; b.tmax = r.tmax; b.pad = sr.pad;
    return true;
}

// Conservative beam / box test; tenter = lower bound of the entry distance of any ray of the beam.
// For every ray r of the beam and every box B' inside [lo, hi]: r's fused slab test accepts B'  =>  this accepts.
// (per axis, after the mirror: near plane l, far plane h; t_near >= (l - omax) * (that >= 0 ? imin : imax) and
//  t_far <= (h - omin) * (that >= 0 ? imax : imin) for every origin / reciprocal inside the bounds; the rounding of these
//  few operations and of the lanes' own arithmetic is covered by the relative 2^-18 and the absolute pad.)
YRT_HD bool beam_test(const Beam& b, const float4& lo, const float4& hi, float& tenter) {
    const float lx = b.sx > 0.f ? lo.x : -hi.x, hx = b.sx > 0.f ? hi.x : -lo.x;
    const float ly = b.sy > 0.f ? lo.y : -hi.y, hy = b.sy > 0.f ? hi.y : -lo.y;
    const float lz = b.sz > 0.f ? lo.z : -hi.z, hz = b.sz > 0.f ? hi.z : -lo.z;
    const float nx = lx - b.omax.x, ny = ly - b.omax.y, nz = lz - b.omax.z;
    const float fx = hx - b.omin.x, fy = hy - b.omin.y, fz = hz - b.omin.z;
    const float tnx = nx * (nx >= 0.f ? b.imin.x : b.imax.x), tny = ny * (ny >= 0.f ? b.imin.y : b.imax.y), tnz = nz * (nz >= 0.f ? b.imin.z : b.imax.z);
    const float tfx = fx * (fx >= 0.f ? b.imax.x : b.imin.x), tfy = fy * (fy >= 0.f ? b.imax.y : b.imin.y), tfz = fz * (fz >= 0.f ? b.imax.z : b.imin.z);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, b.tmin)), tf = fminf(fminf(tfx, tfy), fminf(tfz, b.tmax));
    tenter = tn;
    return !(tn > tf + (fabsf(tn) + fabsf(tf)) * 3.8146973e-6f + b.pad);      // 2^-18; NaN accepts
}

// ---- the wide hierarchy (built from the binary instance tree; kernels in yrt_build.cu, loops in tests/host_emu) ---------
// Level L (L = 0, 1, ...) cuts the binary tree into its maximal subtrees of at most 32^(L+1) instances: "groups" with the
// tight boxes of those subtrees.  The instances below a subtree are consecutive slots and the groups of a level, in slot
// order, partition the slots, so the children of a group are a consecutive run of groups of the level below (level 0: a
// run of at most 32 instance slots).  The top level is the single group of the root.  One array of (lo, hi) float4 pairs:
// first the per-slot boxes, then the levels; lo.w / hi.w of a group = first child / number of children.
struct BeamLevels {
    int levels;                                  // group levels (0 = no instances)
    int threshold[YRT_BEAM_MAX_LEVELS];          // 32^(L+1)
};
inline BeamLevels beam_levels_for(int n_slots) {
    BeamLevels L;
    L.levels = 0;
    long long t = YRT_BEAM_FANOUT;
    for (int k = 0; k < YRT_BEAM_MAX_LEVELS; k++) L.threshold[k] = 0;
    while (n_slots > 0 && L.levels < YRT_BEAM_MAX_LEVELS) {
        L.threshold[L.levels++] = (int)std::min<long long>(t, 0x7fffffff);
        if (t >= n_slots) break;
        t *= YRT_BEAM_FANOUT;
    }
    if (L.levels && L.threshold[L.levels - 1] < n_slots) L.threshold[L.levels - 1] = 0x7fffffff;   // (more than 32^5 instances: the last level takes the rest)
    return L;
}
// box k of the array: slots at [0, n), level L at [(L + 1) n, (L + 2) n) (a level never has more groups than slots)
YRT_HD size_t beam_box_index(int n_slots, int level /* -1 = slots */, int k) { return (size_t)(level + 1) * (size_t)n_slots + (size_t)k; }

YRT_HD float next_down_(float x) { return x - fmaxf(fabsf(x) * 1.1920929e-7f, 1.0e-37f); }   // a float strictly below x (x finite)
YRT_HD float next_up_(float x) { return x + fmaxf(fabsf(x) * 1.1920929e-7f, 1.0e-37f); }

// instance slot k: the (c, h') pair the lanes test (h' inflated exactly like the node boxes, box_center_half) and the beam's
// box [c - h', c + h'] rounded outwards, so that it contains the box the lanes see
YRT_HD void beam_leaf_item(const float4& ilo, const float4& ihi, float4* ch, float4* lohi) {
    float cx, cy, cz, hx, hy, hz;
    box_center_half(ilo.x, ihi.x, cx, hx);
    box_center_half(ilo.y, ihi.y, cy, hy);
    box_center_half(ilo.z, ihi.z, cz, hz);
    ch[0] = mk4(cx, cy, cz, 0.f);
    ch[1] = mk4(hx, hy, hz, 0.f);
    lohi[0] = mk4(next_down_(cx - hx), next_down_(cy - hy), next_down_(cz - hz), 0.f);
    lohi[1] = mk4(next_up_(cx + hx), next_up_(cy + hy), next_up_(cz + hz), 0.f);
}

// what the cut needs from the binary tree over the slots (yrt_lbvh.cuh arrays of the instance tree)
struct BeamTree {
    int n;                         // slots
    const int *parent_int, *parent_leaf, *range_first, *range_last;   // internal nodes 0..n-2; parent of slot k
    const float4 *node_lo, *node_hi;                                    // boxes of the internal nodes (exact unions of the instance boxes)
};
YRT_HD int beam_node_size_(const BeamTree& t, int i) { return t.range_last[i] - t.range_first[i] + 1; }
// does a group of the level with threshold T start at internal node i / at slot k?  flag[first slot] = 1, gnode[first slot] = who
YRT_HD void beam_flag_node_item(const BeamTree& t, int i, int T, int* flag, int* gnode) {
    const int p = t.parent_int[i];
    if (beam_node_size_(t, i) <= T && (p < 0 || beam_node_size_(t, p) > T)) { flag[t.range_first[i]] = 1; gnode[t.range_first[i]] = i; }
}
YRT_HD void beam_flag_slot_item(const BeamTree& t, int k, int T, int* flag, int* gnode) {
    const int p = t.parent_leaf[k];
    if (p < 0 || beam_node_size_(t, p) > T) { flag[k] = 1; gnode[k] = ~k; }
}
// the group that starts at slot k (flag[k] != 0): box of its subtree, widened to contain the inflated boxes the lanes test,
// and its run of children.  gidx / gidx_below: exclusive prefix sums of this level's / the lower level's flags; n_below =
// number of groups of the lower level (level 0: children are slots)
YRT_HD void beam_group_item(const BeamTree& t, int k, int level, const int* gnode, const int* gidx, const int* gidx_below, int n_below, float4* boxes) {
    const int who = gnode[k];
    const int count = who >= 0 ? beam_node_size_(t, who) : 1;
    float4 lo, hi;
    if (who >= 0) { lo = t.node_lo[who]; hi = t.node_hi[who]; }
    else { lo = boxes[2 * beam_box_index(t.n, -1, k)]; hi = boxes[2 * beam_box_index(t.n, -1, k) + 1]; }
    // (1 + 16u) of the lanes' half-extents and the rounding of c, h: a relative 2^-18 of the extent and of the coordinates
    const float ex = (hi.x - lo.x) * 3.8146973e-6f + fmaxf(fabsf(lo.x), fabsf(hi.x)) * 4.7683716e-7f + 1.0e-30f;
    const float ey = (hi.y - lo.y) * 3.8146973e-6f + fmaxf(fabsf(lo.y), fabsf(hi.y)) * 4.7683716e-7f + 1.0e-30f;
    const float ez = (hi.z - lo.z) * 3.8146973e-6f + fmaxf(fabsf(lo.z), fabsf(hi.z)) * 4.7683716e-7f + 1.0e-30f;
    int child_first = k, child_count = count;
    if (level > 0) {
        const int end = k + count;
        child_first = gidx_below[k];
        child_count = (end < t.n ? gidx_below[end] : n_below) - child_first;
    }
    float4* out = boxes + 2 * beam_box_index(t.n, level, gidx[k]);
    out[0] = mk4(lo.x - ex, lo.y - ey, lo.z - ez, int_as_float(child_first));
    out[1] = mk4(hi.x + ex, hi.y + ey, hi.z + ez, int_as_float(child_count));
}

// ---- device: the warp-level walk ---------------------------------------------------------------------------------------
#if defined(__CUDACC__)
// min / max of a float over the whole warp with REDUX (the integer image of a float orders like the float once negative
// values are flipped; `nonneg` skips the flip for values known to be >= 0)
__device__ __forceinline__ float warp_min_f(float v, bool nonneg) {
    int k = __float_as_int(v);
    if (!nonneg) k = k >= 0 ? k : k ^ 0x7fffffff;
    k = __reduce_min_sync(0xffffffffu, k);
    if (!nonneg) k = k >= 0 ? k : k ^ 0x7fffffff;
    return __int_as_float(k);
}
__device__ __forceinline__ float warp_max_f(float v, bool nonneg) {
    int k = __float_as_int(v);
    if (!nonneg) k = k >= 0 ? k : k ^ 0x7fffffff;
    k = __reduce_max_sync(0xffffffffu, k);
    if (!nonneg) k = k >= 0 ? k : k ^ 0x7fffffff;
    return __int_as_float(k);
}

// per-warp scratch in shared memory
struct BeamScratch {
    int stack[2 * YRT_BEAM_STACK];          // (level << 26 | number of children, first child) pairs
    int cand[YRT_BEAM_CAND_CAP];
    float cand_e[YRT_BEAM_CAND_CAP];
};

// bounds of the warp's ray origins (mirrored later, per light / per beam): computed once per warp of shadow rays, whose
// origins are the same for every light; camera rays share one origin and need none of it
struct BeamOrigins { vec3 lo, hi; };
__device__ __forceinline__ BeamOrigins warp_beam_origins(const vec3& o, bool mine) {
    const float big = 3.0e38f;
    BeamOrigins r;
    r.lo = mk3(warp_min_f(mine ? o.x : big, false), warp_min_f(mine ? o.y : big, false), warp_min_f(mine ? o.z : big, false));
    r.hi = mk3(warp_max_f(mine ? o.x : -big, false), warp_max_f(mine ? o.y : -big, false), warp_max_f(mine ? o.z : -big, false));
    return r;
}

// The rays of the lanes in `rays_mask` (each with its slab operands) -> candidate instance slots in sc.cand[0 .. n); with SORT
// in DESCENDING beam entry distance (pushing them in this order leaves the nearest on top of a lane's stack).
// Returns n >= 0, or -1 when the warp must fall back to the per-lane instance tree.  Every lane of the warp must call it
// (lanes outside rays_mask help with the box tests); rays_mask must not be empty.  org = bounds of the rays' origins.
template <bool SORT>
__device__ __forceinline__ int warp_beam_candidates(const SceneView& sv, const ray3& ray, const slabray& sr, const BeamOrigins& org, unsigned rays_mask, int lane,
                                                    BeamScratch& sc, int* visits_out) {
    const unsigned FULL = 0xffffffffu;
    const bool mine = (rays_mask >> lane) & 1u;
    // one vote decides: every ray far enough from the axis planes (the rays that take the reference's slab formula never
    // join a beam) and all reciprocal directions of one sign per axis
    const float ax = fabsf(ray.d.x), ay = fabsf(ray.d.y), az = fabsf(ray.d.z);
    const bool fit = fminf(fminf(ax, ay), az) >= 1.0f / 4096.0f && fmaxf(fmaxf(ax, ay), az) <= 1.0e30f;
    if (__ballot_sync(FULL, mine && !fit) != 0u) return -1;
    const unsigned nx = __ballot_sync(FULL, mine && sr.invd.x < 0.f), ny = __ballot_sync(FULL, mine && sr.invd.y < 0.f), nz = __ballot_sync(FULL, mine && sr.invd.z < 0.f);
    if ((nx != 0u && nx != rays_mask) || (ny != 0u && ny != rays_mask) || (nz != 0u && nz != rays_mask)) return -1;
    Beam b;
    b.sx = nx ? -1.f : 1.f; b.sy = ny ? -1.f : 1.f; b.sz = nz ? -1.f : 1.f;
    // mirrored origin bounds: min / max swap under the mirror
    b.omin = mk3(nx ? -org.hi.x : org.lo.x, ny ? -org.hi.y : org.lo.y, nz ? -org.hi.z : org.lo.z);
    b.omax = mk3(nx ? -org.lo.x : org.hi.x, ny ? -org.lo.y : org.hi.y, nz ? -org.lo.z : org.hi.z);
    const float big = 3.0e38f;
    b.imin = mk3(warp_min_f(mine ? sr.ainv.x : big, true), warp_min_f(mine ? sr.ainv.y : big, true), warp_min_f(mine ? sr.ainv.z : big, true));
    b.imax = mk3(warp_max_f(mine ? sr.ainv.x : 0.f, true), warp_max_f(mine ? sr.ainv.y : 0.f, true), warp_max_f(mine ? sr.ainv.z : 0.f, true));
    b.tmin = warp_min_f(mine ? fmaxf(ray.tmin, 0.f) : big, true);
    b.tmax = warp_max_f(mine ? fmaxf(ray.tmax, 0.f) : 0.f, true);
    b.pad = 4.0f * warp_max_f(mine ? sr.pad : 0.f, true) + 1.0e-30f;

    const unsigned lt = (1u << lane) - 1u;
    const int ns = sv.n_active_instances;
    int sp = 0, n = 0, visits = 0;
    {   // the top level is the single group of the root: its record holds its run of children
        const float4* grp = sv.beam_boxes + 2 * beam_box_index(ns, sv.beam_levels - 1, 0);
        if (lane == 0) { sc.stack[0] = ((sv.beam_levels - 1) << 26) | float_as_int(ld4(grp + 1).w); sc.stack[1] = float_as_int(ld4(grp).w); }
        sp = 1;
    }
    while (sp > 0) {
        __syncwarp();
        sp--;
        const int ent = sc.stack[2 * sp], child_first = sc.stack[2 * sp + 1];
        const int lvl = ent >> 26, child_count = ent & 0x3ffffff;
        const float4* boxes = sv.beam_boxes + 2 * beam_box_index(ns, lvl - 1, child_first);
        for (int c0 = 0; c0 < child_count; c0 += YRT_BEAM_FANOUT) {      // (one round, except for the few groups with more than 32 children)
            bool hit = false;
            float e = 0.f;
            float4 lo = mk4(0.f, 0.f, 0.f, 0.f), hi = lo;
            if (c0 + lane < child_count) {
                lo = ld4(boxes + 2 * (size_t)(c0 + lane));
                hi = ld4(boxes + 2 * (size_t)(c0 + lane) + 1);
                hit = beam_test(b, lo, hi, e);
            }
            visits++;
            const unsigned m = __ballot_sync(FULL, hit);
            const int k = __popc(m);
            if (lvl == 0) {
                if (n + k > YRT_BEAM_CAND_CAP) return -1;
                if (hit) { const int p = n + __popc(m & lt); sc.cand[p] = child_first + c0 + lane; if (SORT) sc.cand_e[p] = e; }
                n += k;
            } else {
                if (sp + k > YRT_BEAM_STACK) return -1;
                __syncwarp();      // every lane has read its entry before a slot is written again
                if (hit) {         // the child's own record carries its run of children: nothing to fetch when it is popped
                    const int p = sp + __popc(m & lt);
                    sc.stack[2 * p] = ((lvl - 1) << 26) | float_as_int(hi.w);
                    sc.stack[2 * p + 1] = float_as_int(lo.w);
                }
                sp += k;
            }
        }
    }
    __syncwarp();
    if (visits_out) *visits_out = visits;
    if (SORT && n > 1) {
        // descending entry distance: lane j ranks candidate j (ties by position), then the list is rewritten in that order
        int slot = 0, rank = 0;
        float e = 0.f;
        if (lane < n) {
            slot = sc.cand[lane]; e = sc.cand_e[lane];
            for (int k = 0; k < n; k++) {
                const float ek = sc.cand_e[k];
                rank += (ek > e || (ek == e && k < lane)) ? 1 : 0;
            }
        }
        __syncwarp();
        if (lane < n) sc.cand[rank] = slot;
        __syncwarp();
    }
    return n;
}
#endif   // __CUDACC__

// host: the same walk with loops in place of the lanes (tests/host_emu).  rays / srs / alive: the warp's up to 32 rays.
// Returns the number of candidates (descending beam entry distance) or -1 = fall back.
inline int host_beam_candidates(const SceneView& sv, const ray3* rays, const slabray* srs, const bool* alive, int n_lanes, int* cand, int* visits_out) {
    int first = -1;
    for (int l = 0; l < n_lanes; l++) if (alive[l]) { first = l; break; }
    if (first < 0) return 0;
    const bool neg_x = srs[first].invd.x < 0.f, neg_y = srs[first].invd.y < 0.f, neg_z = srs[first].invd.z < 0.f;
    Beam b;
    b.sx = neg_x ? -1.f : 1.f; b.sy = neg_y ? -1.f : 1.f; b.sz = neg_z ? -1.f : 1.f;
    b.omin = mk3(FLT_MAX, FLT_MAX, FLT_MAX); b.omax = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    b.imin = mk3(FLT_MAX, FLT_MAX, FLT_MAX); b.imax = mk3(0.f, 0.f, 0.f);
    b.tmin = FLT_MAX; b.tmax = -FLT_MAX; b.pad = 0.f;
    for (int l = 0; l < n_lanes; l++) {
        if (!alive[l]) continue;
        BeamLane bl;
        if (!beam_lane_terms(rays[l], srs[l], neg_x, neg_y, neg_z, bl)) return -1;
        b.omin = mk3(fminf(b.omin.x, bl.o.x), fminf(b.omin.y, bl.o.y), fminf(b.omin.z, bl.o.z));
        b.omax = mk3(fmaxf(b.omax.x, bl.o.x), fmaxf(b.omax.y, bl.o.y), fmaxf(b.omax.z, bl.o.z));
        b.imin = mk3(fminf(b.imin.x, bl.i.x), fminf(b.imin.y, bl.i.y), fminf(b.imin.z, bl.i.z));
        b.imax = mk3(fmaxf(b.imax.x, bl.i.x), fmaxf(b.imax.y, bl.i.y), fmaxf(b.imax.z, bl.i.z));
        b.tmin = fminf(b.tmin, bl.tmin); b.tmax = fmaxf(b.tmax, bl.tmax); b.pad = fmaxf(b.pad, bl.pad);
    }
    b.pad = 4.0f * b.pad + 1.0e-30f;
    int stack[YRT_BEAM_STACK];
    float cand_e[YRT_BEAM_CAND_CAP];
    const int ns = sv.n_active_instances;
    int sp = 0, n = 0, visits = 0;
    stack[sp++] = (sv.beam_levels - 1) << 24;
    while (sp > 0) {
        const int ent = stack[--sp];
        const int lvl = ent >> 24, g = ent & 0xffffff;
        const float4* grp = sv.beam_boxes + 2 * beam_box_index(ns, lvl, g);
        const int child_first = float_as_int(grp[0].w), child_count = float_as_int(grp[1].w);
        for (int c0 = 0; c0 < child_count; c0 += YRT_BEAM_FANOUT) {
            visits++;
            int hits[YRT_BEAM_FANOUT], k = 0;
            float es[YRT_BEAM_FANOUT];
            for (int lane = 0; lane < YRT_BEAM_FANOUT && c0 + lane < child_count; lane++) {
                const int c = child_first + c0 + lane;
                const float4* bx = sv.beam_boxes + 2 * beam_box_index(ns, lvl - 1, c);
                float e;
                if (beam_test(b, bx[0], bx[1], e)) { hits[k] = c; es[k] = e; k++; }
            }
            if (lvl == 0) {
                if (n + k > YRT_BEAM_CAND_CAP) return -1;
                for (int j = 0; j < k; j++) { cand[n + j] = hits[j]; cand_e[n + j] = es[j]; }
                n += k;
            } else {
                if (sp + k > YRT_BEAM_STACK) return -1;
                for (int j = 0; j < k; j++) stack[sp + j] = ((lvl - 1) << 24) | hits[j];
                sp += k;
            }
        }
    }
    if (visits_out) *visits_out = visits;
    // same ranking as the device
    int sorted[YRT_BEAM_CAND_CAP];
    for (int j = 0; j < n; j++) {
        int rank = 0;
        for (int k = 0; k < n; k++) rank += (cand_e[k] > cand_e[j] || (cand_e[k] == cand_e[j] && k < j)) ? 1 : 0;
        sorted[rank] = cand[j];
    }
    for (int j = 0; j < n; j++) cand[j] = sorted[j];
    return n;
}

}  // namespace yrt
