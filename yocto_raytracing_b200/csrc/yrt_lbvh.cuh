// yrt_lbvh.cuh — per-item functions of the GPU LBVH build (Morton keys -> radix sort -> Karras
// topology -> bottom-up refit -> rotations -> traversal nodes (binary or collapsed 4-wide) with small-subtree leaves).
//
// Replaces build_bvh / make_node / split_prims (src/scene.cpp:508-658): the reference builds
// top-down midpoint-split trees serially; topology is free for parity (SURVEY finding 4), the
// primitive bounds are the reference's (scene.cpp:521-546, bbox_to_world src/vmath.h:312-326).
// One build handles MANY trees at once: every item carries a segment id (its shape) in the top
// 16 bits of the 64-bit key, so after the sort each segment is a contiguous range and — because no
// other key shares its prefix — a subtree of the global radix tree; that subtree's root is the
// segment's BLAS root.  The TLAS is the same build with a single segment.
//
// Item functions are __host__ __device__: the kernels in yrt_build.cu call them one thread per
// item, tools/host_emu calls them in serial loops (tests only).
#pragma once
#include "yrt_scene.cuh"

namespace yrt {

#define YRT_LEAF_SIZE_BLAS 3   /* subtrees of <= 3 elements become one leaf (measured 1..5: 17.89 / 16.59 / 16.41 / 16.62 / 16.64 ms; the reference stops at <= 4, scene.cpp:583) */
#define YRT_SIZE_BITS_BLAS 0
#define YRT_SIZE_BITS_TLAS 3
#ifndef YRT_ROTATE_ROUNDS_BLAS
#define YRT_ROTATE_ROUNDS_BLAS 3   /* bottom-up passes of tree rotations after the refit (see rotate_refit_item); measured 0/2/3: 14.41 / 14.03 / .. ms with TLAS 0 */
#endif
#ifndef YRT_ROTATE_ROUNDS_TLAS
#define YRT_ROTATE_ROUNDS_TLAS 2
#endif
#define YRT_ROTATE_PAIRS_BLAS 1     /* rotations also try the other two pairings of the four grandchildren ... */
#define YRT_ROTATE_PAIRS_TLAS 0     /* ... which costs the instance tree more than it gains (host emulation, box tests per ray) */
#define YRT_LEAF_SIZE_TLAS 1   /* one instance per TLAS leaf: its world box is tested before the ray is transformed */

// ---- order-preserving float <-> int for atomic min/max --------------------------------------
YRT_HD int float_to_ordered(float f) { int i = float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
YRT_HD float ordered_to_float(int i) { return int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

#if defined(__CUDA_ARCH__)
#define YRT_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define YRT_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define YRT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define YRT_FENCE() __threadfence()
#else
template <class T> inline T yrt_host_min_(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T yrt_host_max_(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T yrt_host_add_(T* p, T v) { T o = *p; *p = o + v; return o; }
#define YRT_ATOMIC_MIN(p, v) yrt_host_min_((p), (v))
#define YRT_ATOMIC_MAX(p, v) yrt_host_max_((p), (v))
#define YRT_ATOMIC_ADD(p, v) yrt_host_add_((p), (v))
#define YRT_FENCE()
#endif

// ---- flattened geometry as uploaded (what the bounds / gather kernels read) -----------------
struct GeomView {
    const int* shape_kind;
    const int* shape_elem_off;   // into elem_idx (ints)
    const int* shape_elem_cnt;
    const int* shape_vert_off;
    const int* shape_prim_off;   // first global prim id of the shape (exclusive scan of elem_cnt)
    const int* elem_idx;
    const float* pos;            // 3 per vertex
    const float* norm;
    const float* uv;             // 2 per vertex (zeros where the shape has none)
    const float* radius;         // 1 per vertex (zeros where the shape has none)
    const int* prim_shape;       // [n_prims] shape of each global prim id
    int n_prims;
};

YRT_HD vec3 ld3(const float* p, int i) { return mk3(p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]); }

struct Box { vec3 lo, hi; };
YRT_HD Box box_invalid() { Box b; b.lo = mk3(FLT_MAX, FLT_MAX, FLT_MAX); b.hi = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX); return b; }
YRT_HD void box_expand(Box& b, const vec3& lo, const vec3& hi) {   // vmath.h:292-296
    b.lo = mk3(rmin(b.lo.x, lo.x), rmin(b.lo.y, lo.y), rmin(b.lo.z, lo.z));
    b.hi = mk3(rmax(b.hi.x, hi.x), rmax(b.hi.y, hi.y), rmax(b.hi.z, hi.z));
}
// expand_bbox(bbox, p, r) of scene.cpp:521-523
YRT_HD void box_expand_pr(Box& b, const vec3& p, float r) { box_expand(b, p - mk3(r, r, r), p + mk3(r, r, r)); }

// bounds of global prim `g` in its shape's object space (scene.cpp:527-547)
YRT_HD Box prim_bounds(const GeomView& g, int gp) {
    int s = g.prim_shape[gp];
    int e = gp - g.shape_prim_off[s];
    int kind = g.shape_kind[s];
    int vo = g.shape_vert_off[s];
    Box b = box_invalid();
    if (kind == 0) {
        const int* t = g.elem_idx + g.shape_elem_off[s] + 3 * (size_t)e;
        box_expand_pr(b, ld3(g.pos, vo + t[0]), 0.f);
        box_expand_pr(b, ld3(g.pos, vo + t[1]), 0.f);
        box_expand_pr(b, ld3(g.pos, vo + t[2]), 0.f);
    } else if (kind == 1) {
        const int* t = g.elem_idx + g.shape_elem_off[s] + 2 * (size_t)e;
        box_expand_pr(b, ld3(g.pos, vo + t[0]), g.radius[vo + t[0]]);
        box_expand_pr(b, ld3(g.pos, vo + t[1]), g.radius[vo + t[1]]);
    } else {
        const int* t = g.elem_idx + g.shape_elem_off[s] + (size_t)e;
        box_expand_pr(b, ld3(g.pos, vo + t[0]), g.radius[vo + t[0]]);
    }
    return b;
}

// world bounds of an instance: bbox_to_world (vmath.h:312-326) of the shape's root box
YRT_HD Box instance_bounds(const frame3& f, const Box& b) {
    Box w = box_invalid();
    for (int c = 0; c < 8; c++) {
        vec3 p = mk3((c & 4) ? b.hi.x : b.lo.x, (c & 2) ? b.hi.y : b.lo.y, (c & 1) ? b.hi.z : b.lo.z);
        vec3 q = transform_point(f, p);
        box_expand(w, q, q);
    }
    return w;
}

// ---- build items ------------------------------------------------------------------------------
struct LbvhArrays {
    int n;                 // number of items (leaves)
    int n_seg;
    // inputs
    float4* box_lo;        // [n] item bounds (xyz), .w unused
    float4* box_hi;
    const int* seg_of;     // [n] segment of item i (non-decreasing in i)
    const int* seg_first;  // [n_seg+1] first item of each segment
    // per segment reductions (ordered-int encoding): centroid bounds and full bounds
    int* seg_cent_lo;      // [3*n_seg]
    int* seg_cent_hi;
    int* seg_box_lo;       // [3*n_seg]
    int* seg_box_hi;
    // sort
    unsigned long long* keys;   // [n]
    int* order;                 // [n] item id at sorted slot k
    // topology over sorted slots: internal nodes 0..n-2, leaves 0..n-1
    int* left;             // [n-1] child: >=0 internal, <0 => ~leaf slot
    int* right;
    int* range_first;      // [n-1]
    int* range_last;
    int* parent_int;       // [n-1] parent of an internal node (-1 root)
    int* parent_leaf;      // [n]
    int* flags;            // [n-1] refit arrival counters
    int* count;            // [n-1] leaves below an internal node (tree rotations + re-layout)
    int* new_slot;         // [n] sorted slot of each leaf after the re-layout
    int* order_tmp;        // [n]
    int* parent_leaf_tmp;  // [n]
    float4* node_lo;       // [n-1] refit bounds of internal nodes
    float4* node_hi;
    // outputs
    float4* nodes2;        // [4*(n-1)] binary node records
    float4* nodes4;        // [8*(n-1)] 4-wide node records
    int* seg_root;         // [n_seg] root ref per segment
    int* seg_depth;        // [n_seg] max depth (levels of internal nodes) per segment
    int* need2;            // [n-1] traversal stack entries the subtree below a node can occupy (stackneed_item), binary records
    int* need4;            // ... 4-wide records
    int* seg_need2;        // [n_seg] ... below the segment's root
    int* seg_need4;
    int leaf_size;
    int ref_offset;        // added to every internal-node reference (position of this tree set in the shared node array)
    int size_bits;         // 0..3: top bits of the Morton part hold a size class (see morton_item)
    int rotate_pairs;      // rotations also try the two other pairings of the four grandchildren (helps the BLAS, hurts the TLAS)
};

YRT_HD void seg_bounds_init_item(const LbvhArrays& a, int s) {
    for (int c = 0; c < 3; c++) {
        a.seg_cent_lo[3 * s + c] = float_to_ordered(FLT_MAX);
        a.seg_cent_hi[3 * s + c] = float_to_ordered(-FLT_MAX);
        a.seg_box_lo[3 * s + c] = float_to_ordered(FLT_MAX);
        a.seg_box_hi[3 * s + c] = float_to_ordered(-FLT_MAX);
    }
    a.seg_root[s] = YRT_REF_SENTINEL;
    a.seg_depth[s] = 0;
    a.seg_need2[s] = 0;
    a.seg_need4[s] = 0;
}

YRT_HD vec3 box_centroid(const float4& lo, const float4& hi) {   // (bbox.min + bbox.max) / 2, scene.cpp:531
    return mk3((lo.x + hi.x) / 2.0f, (lo.y + hi.y) / 2.0f, (lo.z + hi.z) / 2.0f);
}

YRT_HD void seg_bounds_item(const LbvhArrays& a, int i) {
    int s = a.seg_of[i];
    float4 lo = a.box_lo[i], hi = a.box_hi[i];
    vec3 c = box_centroid(lo, hi);
    YRT_ATOMIC_MIN(&a.seg_cent_lo[3 * s + 0], float_to_ordered(c.x));
    YRT_ATOMIC_MIN(&a.seg_cent_lo[3 * s + 1], float_to_ordered(c.y));
    YRT_ATOMIC_MIN(&a.seg_cent_lo[3 * s + 2], float_to_ordered(c.z));
    YRT_ATOMIC_MAX(&a.seg_cent_hi[3 * s + 0], float_to_ordered(c.x));
    YRT_ATOMIC_MAX(&a.seg_cent_hi[3 * s + 1], float_to_ordered(c.y));
    YRT_ATOMIC_MAX(&a.seg_cent_hi[3 * s + 2], float_to_ordered(c.z));
    YRT_ATOMIC_MIN(&a.seg_box_lo[3 * s + 0], float_to_ordered(lo.x));
    YRT_ATOMIC_MIN(&a.seg_box_lo[3 * s + 1], float_to_ordered(lo.y));
    YRT_ATOMIC_MIN(&a.seg_box_lo[3 * s + 2], float_to_ordered(lo.z));
    YRT_ATOMIC_MAX(&a.seg_box_hi[3 * s + 0], float_to_ordered(hi.x));
    YRT_ATOMIC_MAX(&a.seg_box_hi[3 * s + 1], float_to_ordered(hi.y));
    YRT_ATOMIC_MAX(&a.seg_box_hi[3 * s + 2], float_to_ordered(hi.z));
}

// spread the low 16 bits of v so that two zero bits separate consecutive bits
YRT_HD unsigned long long spread16(unsigned int v) {
    unsigned long long x = v & 0xffffull;
    x = (x | (x << 16)) & 0x0000ff0000ffull;
    x = (x | (x << 8)) & 0x00f00f00f00full;
    x = (x | (x << 4)) & 0x0c30c30c30c3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

YRT_HD unsigned int quant16(float c, float lo, float hi) {
    float ext = hi - lo;
    if (!(ext > 0.f)) return 0u;
    float t = (c - lo) / ext;
    t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    unsigned int q = (unsigned int)(t * 65535.0f);
    return q > 65535u ? 65535u : q;
}

// key = segment (16 bits) | [size class (size_bits)] | Morton code of the centroid inside the segment's
// centroid box (16 bits per axis, 15 when a size class is present).
// Size class: a plain LBVH buries an item that is much larger than its neighbours (the floor among
// 10 000 unit-sized instances) deep in the tree, so every ancestor on that path inherits its huge box
// and every ray has to walk that path.  Putting floor(log2(segment extent / item extent)) in the top
// key bits makes items of very different size separate at the top levels of the radix tree instead
// (cf. extended Morton codes, Vinkler et al. 2017); items of one class form an ordinary spatial LBVH.
YRT_HD void morton_item(const LbvhArrays& a, int i) {
    int s = a.seg_of[i];
    float4 lo = a.box_lo[i], hi = a.box_hi[i];
    vec3 c = box_centroid(lo, hi);
    unsigned int qx = quant16(c.x, ordered_to_float(a.seg_cent_lo[3 * s + 0]), ordered_to_float(a.seg_cent_hi[3 * s + 0]));
    unsigned int qy = quant16(c.y, ordered_to_float(a.seg_cent_lo[3 * s + 1]), ordered_to_float(a.seg_cent_hi[3 * s + 1]));
    unsigned int qz = quant16(c.z, ordered_to_float(a.seg_cent_lo[3 * s + 2]), ordered_to_float(a.seg_cent_hi[3 * s + 2]));
    unsigned long long m;
    if (a.size_bits > 0) {
        float E = fmaxf(fmaxf(ordered_to_float(a.seg_box_hi[3 * s + 0]) - ordered_to_float(a.seg_box_lo[3 * s + 0]),
                              ordered_to_float(a.seg_box_hi[3 * s + 1]) - ordered_to_float(a.seg_box_lo[3 * s + 1])),
                        ordered_to_float(a.seg_box_hi[3 * s + 2]) - ordered_to_float(a.seg_box_lo[3 * s + 2]));
        float e = fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
        int maxc = (1 << a.size_bits) - 1;
        int cls = maxc;
        if (e > 0.f && E > 0.f) {
            // exponent difference = floor(log2(E)) - floor(log2(e)), integer-only and deterministic
            int de = ((float_as_int(E) >> 23) & 0xff) - ((float_as_int(e) >> 23) & 0xff);
            cls = de < 0 ? 0 : (de > maxc ? maxc : de);
        }
        m = (spread16(qx >> 1) << 2) | (spread16(qy >> 1) << 1) | spread16(qz >> 1);   // 15 bits per axis
        m |= (unsigned long long)cls << 45;
    } else {
        m = (spread16(qx) << 2) | (spread16(qy) << 1) | spread16(qz);
    }
    a.keys[i] = ((unsigned long long)(unsigned)s << 48) | m;
    a.order[i] = i;
}

YRT_HD int clz64_(unsigned long long x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
YRT_HD int clz32_(unsigned int x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// length of the common prefix of sorted keys i and j (index breaks ties); -1 outside [0,n)
YRT_HD int delta_(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32_((unsigned)i ^ (unsigned)j);
    return clz64_(a ^ b);
}

// Karras 2012, one internal node per call
YRT_HD void karras_item(const LbvhArrays& a, int i) {
    const unsigned long long* keys = a.keys;
    int n = a.n;
    int d = (delta_(keys, n, i, i + 1) - delta_(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta_(keys, n, i, i - d);
    int lmax = 2;
    while (delta_(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta_(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta_(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta_(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? d : 0);
    int first = i < j ? i : j, last = i < j ? j : i;
    int lc, rc;
    if (first == gamma) { lc = ~gamma; a.parent_leaf[gamma] = i; } else { lc = gamma; a.parent_int[gamma] = i; }
    if (last == gamma + 1) { rc = ~(gamma + 1); a.parent_leaf[gamma + 1] = i; } else { rc = gamma + 1; a.parent_int[gamma + 1] = i; }
    a.left[i] = lc;
    a.right[i] = rc;
    a.range_first[i] = first;
    a.range_last[i] = last;
    if (i == 0) a.parent_int[0] = -1;
    a.flags[i] = 0;
}

YRT_HD void child_box_(const LbvhArrays& a, int c, float4& lo, float4& hi) {
    if (c < 0) { int it = a.order[~c]; lo = a.box_lo[it]; hi = a.box_hi[it]; }
    else { lo = a.node_lo[c]; hi = a.node_hi[c]; }
}

// bottom-up refit, one leaf per call; the second arrival at a node computes its bounds
YRT_HD void refit_item(const LbvhArrays& a, int leaf) {
    int p = a.parent_leaf[leaf];
    while (p >= 0) {
        YRT_FENCE();
        int old = YRT_ATOMIC_ADD(&a.flags[p], 1);
        if (old == 0) return;
        YRT_FENCE();
        float4 l0, h0, l1, h1;
#if defined(__CUDA_ARCH__)
        // children were written by other threads: read through L2, not the (incoherent) L1
        {
            int c0 = a.left[p], c1 = a.right[p];
            if (c0 < 0) { int it = a.order[~c0]; l0 = a.box_lo[it]; h0 = a.box_hi[it]; }
            else { l0 = __ldcg(&a.node_lo[c0]); h0 = __ldcg(&a.node_hi[c0]); }
            if (c1 < 0) { int it = a.order[~c1]; l1 = a.box_lo[it]; h1 = a.box_hi[it]; }
            else { l1 = __ldcg(&a.node_lo[c1]); h1 = __ldcg(&a.node_hi[c1]); }
        }
#else
        child_box_(a, a.left[p], l0, h0);
        child_box_(a, a.right[p], l1, h1);
#endif
        a.node_lo[p] = mk4(rmin(l0.x, l1.x), rmin(l0.y, l1.y), rmin(l0.z, l1.z), 0.f);
        a.node_hi[p] = mk4(rmax(h0.x, h1.x), rmax(h0.y, h1.y), rmax(h0.z, h1.z), 0.f);
        p = a.parent_int[p];
    }
}

// ---- tree rotations (optional pass between refit and emit) ----------------------------------------
// The radix tree splits every range at the spatial median of its Morton prefix, whatever the boxes look like.  One or
// more bottom-up passes of tree rotations (Kensler 2008) repair the worst of it: at node p with children (L, R) and
// R = (RL, RR), exchanging L with RL or RR leaves every subtree but R untouched, so the surface-area cost of the whole
// tree changes by exactly area(R') - area(R); the best of the (up to) four exchanges is applied when it shrinks that
// area.  Same arrival-counter walk as the refit (a node is handled by the second thread to reach it, when both subtrees
// are final), so no two threads ever touch the same nodes.  Rotations never cross a segment (shape) boundary: the leaf
// SET below p is unchanged, only its internal shape.  Afterwards the leaves below a node are no longer contiguous in
// the sorted order, which the leaf references (first, count) need: relayout_* renumbers the leaf slots in tree order.
#if defined(__CUDA_ARCH__)
#define YRT_LDCG(p) __ldcg(p)
#else
#define YRT_LDCG(p) (*(p))
#endif

YRT_HD float half_area_(const float4& lo, const float4& hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
YRT_HD void union_box_(const float4& l0, const float4& h0, const float4& l1, const float4& h1, float4& lo, float4& hi) {
    lo = mk4(rmin(l0.x, l1.x), rmin(l0.y, l1.y), rmin(l0.z, l1.z), 0.f);
    hi = mk4(rmax(h0.x, h1.x), rmax(h0.y, h1.y), rmax(h0.z, h1.z), 0.f);
}
// box / leaf count / parent link of a child reference whose subtree was finished by another thread
YRT_HD void child_box_cg_(const LbvhArrays& a, int c, float4& lo, float4& hi) {
    if (c < 0) { int it = a.order[~c]; lo = a.box_lo[it]; hi = a.box_hi[it]; }
    else { lo = YRT_LDCG(&a.node_lo[c]); hi = YRT_LDCG(&a.node_hi[c]); }
}
YRT_HD int child_count_cg_(const LbvhArrays& a, int c) { return c < 0 ? 1 : YRT_LDCG(&a.count[c]); }
YRT_HD void set_parent_(const LbvhArrays& a, int c, int p) {
    if (c < 0) a.parent_leaf[~c] = p; else a.parent_int[c] = p;
}

YRT_HD void rotate_refit_item(const LbvhArrays& a, int leaf) {
    int p = a.parent_leaf[leaf];
    while (p >= 0) {
        YRT_FENCE();
        int old = YRT_ATOMIC_ADD(&a.flags[p], 1);
        if (old == 0) return;
        YRT_FENCE();
        int c[2] = {YRT_LDCG(&a.left[p]), YRT_LDCG(&a.right[p])};
        float4 lo[2], hi[2];
        child_box_cg_(a, c[0], lo[0], hi[0]);
        child_box_cg_(a, c[1], lo[1], hi[1]);
        // the Karras range of p is still its leaf set (rotations below p only reshaped it)
        bool one_segment = a.seg_of[a.order[a.range_first[p]]] == a.seg_of[a.order[a.range_last[p]]];
        if (one_segment) {
            float best = 0.f;
            int best_side = -1, best_g = 0;
            float4 best_lo = lo[0], best_hi = hi[0];
            for (int side = 0; side < 2; side++) {          // `side` = the internal child that is reshaped
                int r = c[side];                              // (the other child, c[1 - side], is "l" in the comments below)
                if (r < 0) continue;
                int g[2] = {YRT_LDCG(&a.left[r]), YRT_LDCG(&a.right[r])};
                float ar = half_area_(lo[side], hi[side]);
                for (int k = 0; k < 2; k++) {                 // grandchild g[k] moves up, l takes its place next to g[1-k]
                    float4 gl, gh, nl, nh;
                    child_box_cg_(a, g[1 - k], gl, gh);
                    union_box_(lo[1 - side], hi[1 - side], gl, gh, nl, nh);
                    float gain = ar - half_area_(nl, nh);
                    if (gain > best) { best = gain; best_side = side; best_g = k; best_lo = nl; best_hi = nh; }
                }
            }
            // both children internal: the two other pairings of the four grandchildren (LL,LR | RL,RR) ->
            // (LL,RL | LR,RR) and (LL,RR | LR,RL); cost change = area(L') + area(R') - area(L) - area(R)
            int pair = -1;
            float4 pl0, ph0, pl1, ph1;
            if (a.rotate_pairs && c[0] >= 0 && c[1] >= 0) {
                int gl[2] = {YRT_LDCG(&a.left[c[0]]), YRT_LDCG(&a.right[c[0]])};
                int gr[2] = {YRT_LDCG(&a.left[c[1]]), YRT_LDCG(&a.right[c[1]])};
                float4 bl[4], bh[4];
                child_box_cg_(a, gl[0], bl[0], bh[0]); child_box_cg_(a, gl[1], bl[1], bh[1]);
                child_box_cg_(a, gr[0], bl[2], bh[2]); child_box_cg_(a, gr[1], bl[3], bh[3]);
                float cur = half_area_(lo[0], hi[0]) + half_area_(lo[1], hi[1]);
                for (int k = 0; k < 2; k++) {   // LL pairs with RL (k = 0) or RR (k = 1)
                    float4 l0_, h0_, l1_, h1_;
                    union_box_(bl[0], bh[0], bl[2 + k], bh[2 + k], l0_, h0_);
                    union_box_(bl[1], bh[1], bl[3 - k], bh[3 - k], l1_, h1_);
                    float gain = cur - half_area_(l0_, h0_) - half_area_(l1_, h1_);
                    if (gain > best) { best = gain; pair = k; best_side = -1; pl0 = l0_; ph0 = h0_; pl1 = l1_; ph1 = h1_; }
                }
                if (pair >= 0) {
                    // L keeps LL and takes gr[pair]; R keeps gr[1 - pair] and takes LR
                    int L = c[0], R = c[1], take = gr[pair], give = gl[1];
                    a.right[L] = take; set_parent_(a, take, L);
                    if (pair == 0) a.left[R] = give; else a.right[R] = give;
                    set_parent_(a, give, R);
                    a.node_lo[L] = pl0; a.node_hi[L] = ph0;
                    a.node_lo[R] = pl1; a.node_hi[R] = ph1;
                    a.count[L] = child_count_cg_(a, gl[0]) + child_count_cg_(a, take);
                    a.count[R] = child_count_cg_(a, gr[1 - pair]) + child_count_cg_(a, give);
                    lo[0] = pl0; hi[0] = ph0; lo[1] = pl1; hi[1] = ph1;
                }
            }
            if (best_side >= 0) {
                int r = c[best_side], l = c[1 - best_side];
                int g[2] = {YRT_LDCG(&a.left[r]), YRT_LDCG(&a.right[r])};
                int up = g[best_g], stay = g[1 - best_g];
                // r keeps `stay` and adopts l; p keeps r and adopts `up`
                if (best_g == 0) a.left[r] = l; else a.right[r] = l;
                a.node_lo[r] = best_lo;
                a.node_hi[r] = best_hi;
                a.count[r] = child_count_cg_(a, l) + child_count_cg_(a, stay);
                set_parent_(a, l, r);
                if (best_side == 0) a.right[p] = up; else a.left[p] = up;
                set_parent_(a, up, p);
                // refresh p's view of its children
                c[1 - best_side] = up;
                child_box_cg_(a, up, lo[1 - best_side], hi[1 - best_side]);
                lo[best_side] = best_lo; hi[best_side] = best_hi;
            }
        }
        float4 nl, nh;
        union_box_(lo[0], hi[0], lo[1], hi[1], nl, nh);
        a.node_lo[p] = nl;
        a.node_hi[p] = nh;
        a.count[p] = child_count_cg_(a, c[0]) + child_count_cg_(a, c[1]);
        p = YRT_LDCG(&a.parent_int[p]);
    }
}

// new sorted slot of a leaf = number of leaves left of it in tree order; also the new [first, last] of every
// internal node (written by its leftmost / rightmost leaf).  One leaf per call, two walks to the root.
YRT_HD void relayout_slot_item(const LbvhArrays& a, int leaf) {
    int pos = 0;
    int child = ~leaf;
    for (int p = a.parent_leaf[leaf]; p >= 0; p = a.parent_int[p]) {
        if (a.right[p] == child) { int l = a.left[p]; pos += l < 0 ? 1 : a.count[l]; }
        child = p;
    }
    a.new_slot[leaf] = pos;
    bool leftmost = true, rightmost = true;
    child = ~leaf;
    for (int p = a.parent_leaf[leaf]; p >= 0 && (leftmost || rightmost); p = a.parent_int[p]) {
        if (a.right[p] == child) leftmost = false; else rightmost = false;
        if (leftmost) a.range_first[p] = pos;
        if (rightmost) a.range_last[p] = pos;
        child = p;
    }
}
YRT_HD void relayout_move_item(const LbvhArrays& a, int leaf) {
    int s = a.new_slot[leaf];
    a.order_tmp[s] = a.order[leaf];
    a.parent_leaf_tmp[s] = a.parent_leaf[leaf];
}
YRT_HD void relayout_copy_item(const LbvhArrays& a, int slot) {
    a.order[slot] = a.order_tmp[slot];
    a.parent_leaf[slot] = a.parent_leaf_tmp[slot];
}
YRT_HD void relayout_refs_item(const LbvhArrays& a, int i) {
    int l = a.left[i], r = a.right[i];
    if (l < 0) a.left[i] = ~a.new_slot[~l];
    if (r < 0) a.right[i] = ~a.new_slot[~r];
}

// reference of a child for traversal: subtrees of <= leaf_size items become one leaf
YRT_HD int child_ref_(const LbvhArrays& a, int c) {
    if (c < 0) return make_leaf_ref(~c, 1);
    int cnt = a.range_last[c] - a.range_first[c] + 1;
    if (cnt <= a.leaf_size) return make_leaf_ref(a.range_first[c], cnt);
    return c + a.ref_offset;
}

// true if binary child reference c becomes a leaf of the traversal tree (a single item or a subtree of <= leaf_size items)
YRT_HD bool is_leaf_child_(const LbvhArrays& a, int c) {
    return c < 0 || a.range_last[c] - a.range_first[c] + 1 <= a.leaf_size;
}

// Emit the two traversal records of binary internal node i and, if i is exactly a segment, its root ref.
// Binary record: the two children.  4-wide record: the slots start as the two binary children; while a slot is free, the
// internal (non-leaf) slot with the largest surface area is replaced by its own two children (the usual greedy binary ->
// wide collapse: the box most likely to be entered is the one opened).  Every binary internal node gets both records at
// its own index, so segment roots and child references need no renumbering; of the wide records only those reachable
// from a root through wide links are ever read (about one in three), and each is exactly one 128-byte line.
YRT_HD void emit_item(const LbvhArrays& a, int i) {
    int c[4];
    float4 lo[4], hi[4];
    nodebox bx[4];
    int ref[4];
    c[0] = a.left[i]; c[1] = a.right[i];
    child_box_(a, c[0], lo[0], hi[0]);
    child_box_(a, c[1], lo[1], hi[1]);
    auto finish = [&](int k) {
        box_center_half(lo[k].x, hi[k].x, bx[k].cx, bx[k].hx);
        box_center_half(lo[k].y, hi[k].y, bx[k].cy, bx[k].hy);
        box_center_half(lo[k].z, hi[k].z, bx[k].cz, bx[k].hz);
        ref[k] = child_ref_(a, c[k]);
    };
    finish(0); finish(1);
    node_pack<2>(a.nodes2 + YRT_NODE_STRIDE(2) * (size_t)i, bx, ref, 2);
    int nb = 2;
    while (nb < 4) {
        int best = -1;
        float best_area = -1.f;
        for (int k = 0; k < nb; k++) {
            if (is_leaf_child_(a, c[k])) continue;
            float ar = half_area_(lo[k], hi[k]);
            if (ar > best_area) { best_area = ar; best = k; }
        }
        if (best < 0) break;
        // the opened slot keeps its position for its left child; the right child goes right behind it (siblings stay
        // neighbours, so that pairs (0,1) / (2,3) of a fully opened node are the two binary subtrees)
        int l = a.left[c[best]], r = a.right[c[best]];
        for (int k = nb; k > best + 1; k--) { c[k] = c[k - 1]; lo[k] = lo[k - 1]; hi[k] = hi[k - 1]; }
        c[best] = l; c[best + 1] = r;
        child_box_(a, l, lo[best], hi[best]);
        child_box_(a, r, lo[best + 1], hi[best + 1]);
        nb++;
    }
    for (int k = 0; k < nb; k++) finish(k);
    node_pack<4>(a.nodes4 + YRT_NODE_STRIDE(4) * (size_t)i, bx, ref, nb);
    int first = a.range_first[i], last = a.range_last[i];
    int s = a.seg_of[a.order[first]];
    if (s == a.seg_of[a.order[last]] && first == a.seg_first[s] && last == a.seg_first[s + 1] - 1) {
        a.seg_root[s] = child_ref_(a, i);
    }
}

// segments of exactly one item have no internal node: their root is that leaf (one per segment)
YRT_HD void single_root_item(const LbvhArrays& a, int s) {
    int cnt = a.seg_first[s + 1] - a.seg_first[s];
    if (cnt == 1) a.seg_root[s] = make_leaf_ref(a.seg_first[s], 1);
}

// depth of sorted leaf `leaf` below its segment root (counts internal nodes on the path)
YRT_HD void depth_item(const LbvhArrays& a, int leaf) {
    int s = a.seg_of[a.order[leaf]];
    int sf = a.seg_first[s], sl = a.seg_first[s + 1] - 1;
    if (sl == sf) return;
    int d = 0;
    int p = a.parent_leaf[leaf];
    while (p >= 0) {
        d++;
        if (a.range_first[p] == sf && a.range_last[p] == sl) break;
        p = a.parent_int[p];
    }
    YRT_ATOMIC_MAX(&a.seg_depth[s], d);
}

// Traversal-stack entries the subtree below node record i can occupy at once: a visit leaves at most (children - 1)
// siblings on the stack while it descends into one child, so need(i) = children(i) - 1 + max over internal children —
// for the binary and for the 4-wide records.  Bottom-up with the arrival counters of the refit (flags zeroed before): a
// node is handled by the second thread to reach it, when every record below it has its value.  Runs after emit.
template <int W>
YRT_HD int stackneed_of_(const LbvhArrays& a, int p, const float4* nodes, const int* need) {
    const float4* n = nodes + YRT_NODE_STRIDE(W) * (size_t)p;
    int nb = 0, deepest = 0;
    for (int k = 0; k < W; k++) {
        nodebox b;
        int ref;
        node_child<W>(n, k, b, ref);
        if (!(b.hx >= 0.f)) continue;    // empty slot
        nb++;
        if (ref >= 0) { int c = YRT_LDCG(&need[ref - a.ref_offset]); if (c > deepest) deepest = c; }
    }
    return nb - 1 + deepest;
}
YRT_HD void stackneed_item(const LbvhArrays& a, int leaf) {
    int p = a.parent_leaf[leaf];
    while (p >= 0) {
        YRT_FENCE();
        int old = YRT_ATOMIC_ADD(&a.flags[p], 1);
        if (old == 0) return;
        YRT_FENCE();
        int n2 = stackneed_of_<2>(a, p, a.nodes2, a.need2), n4 = stackneed_of_<4>(a, p, a.nodes4, a.need4);
        a.need2[p] = n2;
        a.need4[p] = n4;
        int first = a.range_first[p], last = a.range_last[p];
        int s = a.seg_of[a.order[first]];
        if (s == a.seg_of[a.order[last]] && first == a.seg_first[s] && last == a.seg_first[s + 1] - 1) { a.seg_need2[s] = n2; a.seg_need4[s] = n4; }
        p = YRT_LDCG(&a.parent_int[p]);
    }
}

}  // namespace yrt
