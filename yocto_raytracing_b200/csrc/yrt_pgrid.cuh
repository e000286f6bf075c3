// yrt_pgrid.cuh — apex grids: the instance level decided per cell of directions, for rays that all pass through one point.
//
// Every camera ray of raytrace() starts at the pinhole (eval_camera, src/raytrace.cpp:6-37: the reference ignores the
// aperture), and every shadow ray of shade() ends at a point light (src/raytrace.cpp:126-133).  For such a bundle most of
// the instance level of the two-level BVH (intersect_bvh(scene…), src/scene.cpp:446-479) can be decided once per
// direction instead of once per ray: the directions through the apex are cut into cells (pixel tiles of the image plane
// for the camera, the six faces of a cube around a light), and every cell gets the list of the instances whose world box a
// ray of that cell can touch, nearest first, each with the part of its box the cell's pyramid can reach.  Two forms:
//   * list form (camera grid): cells[n_cells] = (first entry, count), entries = (clipped box | instance slot | distance
//     bound); a ray walks its cell's entries in order (Tracer<…, GRID>, yrt_trace.cuh), tests each box with the usual slab
//     test, enters the instances it touches, and stops at the first entry whose bound lies beyond the hit it has.  A warp
//     is two neighbouring pixels, i.e. one cell: the walk is coherent;
//   * chain form (light grids): the list is written out as a short chain of ordinary traversal nodes (instance leaves +
//     "the rest of the list") behind the trees in the scene's node array, roots[n_cells] holds where a ray of the cell
//     starts (YRT_REF_DONE for an empty cell) — the traversal code is unchanged, the ray just does not start at the
//     instance tree's root, and chain visits share the instruction stream with the shape-tree visits of other lanes.
// Either way a ray does 2-4 box tests at the instance level instead of 25-35.  What it finds is unchanged: a cell's list
// is a superset of the instances the full tree walk would reach for any of its rays (conservative margins below; audited
// on the host emulation against the tree walk), the primitive tests are the same, ties are settled by the rank tables,
// and rays or cells the grids cannot serve (too many candidates, no room, near-axis-parallel directions, rotated lights,
// mirror rays) start at the tree's root as before.
//
// Built on the device.  Light grids, with the scene: count (one warp per instance, lanes over the cells its projection
// covers; instances covering more than YRT_PGRID_BIG cells by a whole grid of threads) -> allocate (one atomic per warp)
// -> fill -> per cell: insertion sort by distance bound, node chain.  Camera grid, EVERY frame (the camera is a per-call
// argument): one scatter pass into 16 fixed slots per cell, then one thread per slot writes its entry at its rank.
//
// Item functions are __host__ __device__ like the LBVH's: kernels call them per thread, tests/host_emu in serial loops.
#pragma once
#include "yrt_lbvh.cuh"

namespace yrt {

#ifndef YRT_PGRID_MAX_LIST
#define YRT_PGRID_MAX_LIST 32    /* cells with more candidates keep the instance tree (a chain can hold that many postponed entries on the ray's stack) */
#endif
#define YRT_MAX_LIGHT_GRIDS 8    /* lights beyond the first 8 keep the instance tree */
#define YRT_PGRID_BIG 512        /* an instance that covers more cells than this is scattered by a whole grid of threads, not by one warp */

// what a traversal kernel needs to find a ray's starting node (kernel parameter: lives in the constant bank)
struct GridRef {
    const int* roots;       // chain form (light grids): per cell the node reference to start from
    const int2* cells;      // list form (camera grid): per cell (first entry, count); count < 0: the cell is not served
    const float4* entries;  // list form: 2 per entry = (box centre | instance slot), (inflated half-extent | distance bound)
    int nx;                 // cells per row (camera) / per face edge (cube); 0: no grid, the rays start at the instance tree's root
    int shift;              // camera: log2(pixels per cell edge)
};
struct LightGrids { GridRef g[YRT_MAX_LIGHT_GRIDS]; };

YRT_HD GridRef gridref_none() { GridRef g; g.roots = nullptr; g.cells = nullptr; g.entries = nullptr; g.nx = 0; g.shift = 0; return g; }

// build-time description of one grid
struct PGridDesc {
    int mode;              // 1: image plane of a pinhole camera, 2: cube around a point
    int nx, ny;            // cells per face
    int n_cells;           // nx * ny (camera), 6 * nx * ny (cube)
    vec3 apex;
    vec3 ex, ey, ez;       // camera: pixel coordinates of X are (cx + sx a / g, cy + sy b / g) with (a, b, g) = (X - apex) . (ex, ey, ez)
    float sx, sy, cx, cy;
    float inv_cell;        // camera: 1 / (pixels per cell edge)
    float margin;          // camera: pixels; cube: units of the face coordinate u in [-1, 1]
    float near_all;        // a box nearer to the apex than this is a candidate of every cell
    float skip_within;     // a box that lies entirely within this distance of the apex cannot be reached by any ray of the bundle
                           // (shadow rays stop 0.01 short of the light, src/raytrace.cpp:131) and is left out
    float dist_scale, dist_bias;   // stored bound = distance * dist_scale - dist_bias (never above the distance any ray measures)
    int max_list, capacity;
};

struct PGridArrays {
    PGridDesc d;
    const float4* inst_box;        // 2 per instance slot: centre, inflated half-extent (box_center_half)
    int n_inst;
    int* cnt;                      // [n_cells] candidates per cell, then the fill cursors
    int* total;                    // list entries handed out so far
    int2* cells;                   // [n_cells] (first key, count) of each cell's list; count < 0: not served
    unsigned long long* keys;      // [capacity] (distance bound bits << 32 | slot), unsorted
    int* big;                      // [1 + n_inst] number and slots of the instances that cover more than YRT_PGRID_BIG cells
    // output, list form (as_list != 0): cells + entries, the candidates of a cell in order, each with its own box
    int as_list;
    int fixed_k;                   // list form only, > 0: every cell owns fixed_k consecutive keys / entries (first = cell * fixed_k), so one
                                   // scatter pass suffices (no count -> allocate -> fill); a cell with more candidates is not served
    float4* entries;               // [2 * capacity]
    // output, chain form
    int* roots;                    // [n_cells]
    float4* nodes;                 // the scene's node array of arity `arity` (record i at nodes + stride * i)
    int arity;                     // 2 or 4
    int node_first, node_capacity; // the chain nodes of this grid are records [node_first, node_first + node_capacity)
    int* node_total;               // chain nodes handed out so far (may be shared by several grids)
    int tlas_root;                 // where the rays of a cell that is not served start ...
    const int* tlas_root_dev;      // ... or, if not null, where to read it (scene build: the root is not on the host yet)
};
YRT_HD int pgrid_tlas_root(const PGridArrays& a) { return a.tlas_root_dev ? *a.tlas_root_dev : a.tlas_root; }

struct PRect { int x0, y0, x1, y1; };   // inclusive cell ranges

YRT_HD int pgrid_faces(const PGridDesc& d) { return d.mode == 1 ? 1 : 6; }

YRT_HD float comp(const vec3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// Euclidean distance from the apex to the box
YRT_HD float pgrid_box_distance(const PGridDesc& d, const vec3& lo, const vec3& hi) {
    float dx = fmaxf(fmaxf(lo.x - d.apex.x, d.apex.x - hi.x), 0.f);
    float dy = fmaxf(fmaxf(lo.y - d.apex.y, d.apex.y - hi.y), 0.f);
    float dz = fmaxf(fmaxf(lo.z - d.apex.z, d.apex.z - hi.z), 0.f);
    return sqrtf(dx * dx + dy * dy + dz * dz);
}

// cell range [i0, i1] of the interval [a, b] (already in cell units), clamped to [0, n - 1]; false: outside the grid
YRT_HD bool pgrid_range(float a, float b, int n, int& i0, int& i1) {
    if (!(a <= b)) { i0 = 0; i1 = n - 1; return true; }      // NaN: be safe, cover everything
    if (a >= (float)n || b < 0.f) return false;
    float fa = floorf(fmaxf(a, 0.f)), fb = floorf(fminf(b, (float)(n - 1)));
    i0 = (int)fa; i1 = (int)fb;
    return i0 <= i1;
}

// cells of face `face` that rays towards the box [lo, hi] can fall into
YRT_HD bool pgrid_rect(const PGridDesc& d, const vec3& lo, const vec3& hi, int face, PRect& r) {
    if (d.mode == 1) {
        float x0 = FLT_MAX, x1 = -FLT_MAX, y0 = FLT_MAX, y1 = -FLT_MAX;
        bool behind = false;
        for (int c = 0; c < 8; c++) {
            vec3 q = mk3((c & 4) ? hi.x : lo.x, (c & 2) ? hi.y : lo.y, (c & 1) ? hi.z : lo.z) - d.apex;
            float a = dot(q, d.ex), b = dot(q, d.ey), g = dot(q, d.ez);
            if (!(g > d.near_all)) { behind = true; continue; }
            float px = d.cx + d.sx * (a / g), py = d.cy + d.sy * (b / g);
            x0 = fminf(x0, px); x1 = fmaxf(x1, px); y0 = fminf(y0, py); y1 = fmaxf(y1, py);
        }
        if (behind) {
            // the box reaches behind (or next to) the pinhole: its projection is unbounded; if no corner is in front it
            // cannot be seen at all, otherwise it is a candidate everywhere
            if (x0 > x1) return false;
            r.x0 = 0; r.y0 = 0; r.x1 = d.nx - 1; r.y1 = d.ny - 1;
            return true;
        }
        return pgrid_range((x0 - d.margin) * d.inv_cell, (x1 + d.margin) * d.inv_cell, d.nx, r.x0, r.x1) &&
               pgrid_range((y0 - d.margin) * d.inv_cell, (y1 + d.margin) * d.inv_cell, d.ny, r.y0, r.y1);
    }
    // cube face: major axis m, sign s; depth g = s (X - apex)[m] > 0, face coordinates (u, v) = (X - apex)[a1, a2] / g
    const int m = face >> 1, a1 = m == 2 ? 0 : m + 1, a2 = a1 == 2 ? 0 : a1 + 1;
    const bool neg = (face & 1) != 0;
    const vec3 rl = lo - d.apex, rh = hi - d.apex;
    const float g0 = neg ? -comp(rh, m) : comp(rl, m), g1 = neg ? -comp(rl, m) : comp(rh, m);
    if (!(g1 > 0.f)) return false;
    const float inf = int_as_float(0x7f800000);
    float lim = 1.0f + d.margin, half = 0.5f * (float)d.nx;
    float uv0[2], uv1[2];
    for (int k = 0; k < 2; k++) {
        const int ax = k == 0 ? a1 : a2;
        const float x0 = comp(rl, ax), x1 = comp(rh, ax);
        float umin = x0 >= 0.f ? x0 / g1 : (g0 > 0.f ? x0 / g0 : -inf);
        float umax = x1 <= 0.f ? x1 / g1 : (g0 > 0.f ? x1 / g0 : inf);
        if (umin > lim || umax < -lim) return false;
        uv0[k] = (fmaxf(umin, -lim) - d.margin + 1.0f) * half;
        uv1[k] = (fminf(umax, lim) + d.margin + 1.0f) * half;
    }
    return pgrid_range(uv0[0], uv1[0], d.nx, r.x0, r.x1) && pgrid_range(uv0[1], uv1[1], d.ny, r.y0, r.y1);
}

// cell of the direction q (pointing away from the apex) in a cube grid of R x R cells per face
YRT_HD int pgrid_cube_cell(const vec3& q, int R) {
    const float ax = fabsf(q.x), ay = fabsf(q.y), az = fabsf(q.z);
    int m = 0;
    float g = ax, c1 = q.y, c2 = q.z, sm = q.x;
    if (ay > g) { m = 1; g = ay; c1 = q.z; c2 = q.x; sm = q.y; }
    if (az > g) { m = 2; g = az; c1 = q.x; c2 = q.y; sm = q.z; }
    const int face = 2 * m + (sm < 0.f ? 1 : 0);
    const float ig = 1.0f / g, half = 0.5f * (float)R;
    float fu = (c1 * ig + 1.0f) * half, fv = (c2 * ig + 1.0f) * half;
    int iu = (int)fminf(fmaxf(fu, 0.f), (float)(R - 1)), iv = (int)fminf(fmaxf(fv, 0.f), (float)(R - 1));
    return (face * R + iv) * R + iu;
}

// The rays of one cell: apex + s q, s >= 0, q = alpha E1 + beta E2 + E3 with (alpha, beta) anywhere in [a0, a1] x [b0, b1]
// (the cell, margins included).  q_k is affine in (alpha, beta), so per axis it ranges over [qmin_k, qmax_k] taken from the
// cell's corners.
struct PCellDirs { vec3 qmin, qmax; float len_min; };

YRT_HD PCellDirs pgrid_cell_dirs(const PGridDesc& d, int face, int cx, int cy) {
    vec3 E1, E2, E3;
    float a0, a1, b0, b1;
    if (d.mode == 1) {
        E1 = d.ex; E2 = d.ey; E3 = d.ez;
        const float cell = 1.0f / d.inv_cell;
        a0 = ((float)cx * cell - d.margin - d.cx) / d.sx; a1 = ((float)(cx + 1) * cell + d.margin - d.cx) / d.sx;
        b0 = ((float)cy * cell - d.margin - d.cy) / d.sy; b1 = ((float)(cy + 1) * cell + d.margin - d.cy) / d.sy;
    } else {
        const int m = face >> 1, x1 = m == 2 ? 0 : m + 1, x2 = x1 == 2 ? 0 : x1 + 1;
        E1 = mk3(x1 == 0, x1 == 1, x1 == 2); E2 = mk3(x2 == 0, x2 == 1, x2 == 2);
        const float sg = (face & 1) ? -1.0f : 1.0f;
        E3 = mk3(m == 0 ? sg : 0.f, m == 1 ? sg : 0.f, m == 2 ? sg : 0.f);
        const float step = 2.0f / (float)d.nx;
        a0 = step * (float)cx - 1.0f - d.margin; a1 = step * (float)(cx + 1) - 1.0f + d.margin;
        b0 = step * (float)cy - 1.0f - d.margin; b1 = step * (float)(cy + 1) - 1.0f + d.margin;
    }
    PCellDirs c;
    c.qmin = mk3(E3.x + fminf(a0 * E1.x, a1 * E1.x) + fminf(b0 * E2.x, b1 * E2.x), E3.y + fminf(a0 * E1.y, a1 * E1.y) + fminf(b0 * E2.y, b1 * E2.y),
                 E3.z + fminf(a0 * E1.z, a1 * E1.z) + fminf(b0 * E2.z, b1 * E2.z));
    c.qmax = mk3(E3.x + fmaxf(a0 * E1.x, a1 * E1.x) + fmaxf(b0 * E2.x, b1 * E2.x), E3.y + fmaxf(a0 * E1.y, a1 * E1.y) + fmaxf(b0 * E2.y, b1 * E2.y),
                 E3.z + fmaxf(a0 * E1.z, a1 * E1.z) + fmaxf(b0 * E2.z, b1 * E2.z));
    const float am = (a0 <= 0.f && a1 >= 0.f) ? 0.f : fminf(fabsf(a0), fabsf(a1));
    const float bm = (b0 <= 0.f && b1 >= 0.f) ? 0.f : fminf(fabsf(b0), fabsf(b1));
    c.len_min = sqrtf(am * am * dot(E1, E1) + bm * bm * dot(E2, E2) + dot(E3, E3)) * (1.0f - 1.0e-4f);
    return c;
}

// Range [s_lo, s_hi] of the ray parameter over which a ray of the cell can be inside the box [rl, rh] (relative to the
// apex).  Per axis k the slab is entered at rl_k / q_k (q_k > 0) or rh_k / q_k (q_k < 0) and left at the other one; the
// smallest entry / largest exit over q_k in [qmin_k, qmax_k] is at an end of that range (no bound if it contains 0).
// s_lo = the largest per-axis entry minimum bounds every ray's entry from below, s_hi = the smallest per-axis exit
// maximum bounds every exit from above.
YRT_HD void pgrid_cell_span(const PCellDirs& c, const vec3& rl, const vec3& rh, float& s_lo, float& s_hi) {
    s_lo = 0.f;
    s_hi = int_as_float(0x7f800000);
    for (int k = 0; k < 3; k++) {
        const float qmin = comp(c.qmin, k), qmax = comp(c.qmax, k), lo = comp(rl, k), hi = comp(rh, k);
        if (qmin > 0.f) {
            if (lo >= 0.f) s_lo = fmaxf(s_lo, lo / qmax);      // (a negative entry bounds nothing: the ray starts inside the slab)
            s_hi = fminf(s_hi, hi >= 0.f ? hi / qmin : 0.f);   // (hi < 0: the slab lies behind every ray of the cell)
        } else if (qmax < 0.f) {
            if (hi <= 0.f) s_lo = fmaxf(s_lo, hi / qmin);
            s_hi = fminf(s_hi, lo <= 0.f ? lo / qmax : 0.f);
        }
    }
}

// The part of the box a ray of the cell can reach: the box cut down to the bounding box of the cell's pyramid between
// s_lo and s_hi (per axis the coordinates s q_k, the two ranges taken independently: a superset).  A ray of the cell that
// hits something of the instance hits it inside this smaller box, so testing it instead of the whole instance box loses
// nothing — and the boxes of "the rest of the list" stay tight where an instance is much larger than a cell (the floor).
YRT_HD void pgrid_cell_clip(const PGridDesc& d, const PCellDirs& c, const vec3& rl, const vec3& rh, vec3& clo, vec3& chi) {
    clo = rl; chi = rh;
    float s_lo, s_hi;
    pgrid_cell_span(c, rl, rh, s_lo, s_hi);
    s_lo *= 1.0f - 1.0e-3f;
    s_hi *= 1.0f + 1.0e-3f;
    if (!(s_hi < 1.0e30f) || !(s_lo <= s_hi)) return;       // unbounded, or (numerically) unreachable: keep the whole box
    const float pad = d.dist_bias;
    vec3 lo, hi;
    lo = mk3(fminf(s_lo * c.qmin.x, s_hi * c.qmin.x) - pad, fminf(s_lo * c.qmin.y, s_hi * c.qmin.y) - pad, fminf(s_lo * c.qmin.z, s_hi * c.qmin.z) - pad);
    hi = mk3(fmaxf(s_lo * c.qmax.x, s_hi * c.qmax.x) + pad, fmaxf(s_lo * c.qmax.y, s_hi * c.qmax.y) + pad, fmaxf(s_lo * c.qmax.z, s_hi * c.qmax.z) + pad);
    lo = mk3(fmaxf(lo.x, rl.x), fmaxf(lo.y, rl.y), fmaxf(lo.z, rl.z));
    hi = mk3(fminf(hi.x, rh.x), fminf(hi.y, rh.y), fminf(hi.z, rh.z));
    if (!(lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z)) return;
    clo = lo; chi = hi;
}

// largest distance from the apex to a point of the box
YRT_HD float pgrid_box_far_distance(const PGridDesc& d, const vec3& lo, const vec3& hi) {
    float dx = fmaxf(fabsf(lo.x - d.apex.x), fabsf(hi.x - d.apex.x));
    float dy = fmaxf(fabsf(lo.y - d.apex.y), fabsf(hi.y - d.apex.y));
    float dz = fmaxf(fabsf(lo.z - d.apex.z), fabsf(hi.z - d.apex.z));
    return sqrtf(dx * dx + dy * dy + dz * dz);
}

// One (instance slot, lane) pair of the count / fill passes: the lanes stride over the cells the instance covers.
// First pass (big_pass = false, one warp per instance): an instance that covers more than YRT_PGRID_BIG cells is only
// noted in a.big (count pass) and left to the second pass, where all threads of a fixed grid stride over its cells.
YRT_HD void pgrid_scatter_item(const PGridArrays& a, int slot, int lane, int n_lanes, bool fill, bool big_pass) {
    const PGridDesc& d = a.d;
    const float4 bc = a.inst_box[2 * (size_t)slot], bh = a.inst_box[2 * (size_t)slot + 1];
    const vec3 lo = mk3(bc.x - bh.x, bc.y - bh.y, bc.z - bh.z), hi = mk3(bc.x + bh.x, bc.y + bh.y, bc.z + bh.z);
    if (pgrid_box_far_distance(d, lo, hi) < d.skip_within) return;
    const float dist = pgrid_box_distance(d, lo, hi);
    const bool everywhere = !(dist >= d.near_all);
    const vec3 rl = lo - d.apex, rh = hi - d.apex;
    const int nf = pgrid_faces(d), per_face = d.nx * d.ny;
    PRect rect[6];
    int total = 0;
    for (int f = 0; f < nf; f++) {
        PRect& r = rect[f];
        if (everywhere) { r.x0 = 0; r.y0 = 0; r.x1 = d.nx - 1; r.y1 = d.ny - 1; }
        else if (!pgrid_rect(d, lo, hi, f, r)) { r.x0 = 0; r.x1 = -1; r.y0 = 0; r.y1 = -1; }
        total += (r.x1 - r.x0 + 1) * (r.y1 - r.y0 + 1);
    }
    if (!big_pass && total > YRT_PGRID_BIG) {
        if (!fill && lane == 0) a.big[1 + YRT_ATOMIC_ADD(&a.big[0], 1)] = slot;
        return;
    }
    for (int f = 0; f < nf; f++) {
        const PRect r = rect[f];
        const int w = r.x1 - r.x0 + 1, n = w * (r.y1 - r.y0 + 1);
        for (int t = lane; t < n; t += n_lanes) {
            const int cy = r.y0 + t / w, cx = r.x0 + (t - (t / w) * w);
            const int cell = f * per_face + cy * d.nx + cx;
            if (!fill && a.fixed_k <= 0) {
                YRT_ATOMIC_ADD(&a.cnt[cell], 1);
            } else {
                int2 c;
                if (a.fixed_k > 0) { c.x = cell * a.fixed_k; c.y = a.fixed_k; }
                else c = a.cells[cell];
                if (c.y > 0) {
                    // the entry's distance bound: the box's distance from the apex, or — tighter for large boxes — the
                    // nearest point at which a ray of THIS cell can enter it; a hair below either (dist_scale, dist_bias)
                    const PCellDirs cd = pgrid_cell_dirs(d, f, cx, cy);
                    float s_lo, s_hi;
                    pgrid_cell_span(cd, rl, rh, s_lo, s_hi);
                    const float near = fmaxf(dist, s_lo * cd.len_min);
                    const float bound = fmaxf(near * d.dist_scale - d.dist_bias, 0.f);
                    const int pos = YRT_ATOMIC_ADD(&a.cnt[cell], 1);
                    if (pos < c.y) a.keys[(size_t)c.x + pos] = ((unsigned long long)(unsigned)float_as_int(bound) << 32) | (unsigned)slot;
                }
            }
        }
    }
}

// one cell: hand out its run of entries (or send its rays back to the instance tree).  pgrid_alloc_need = entries the cell
// asks for, pgrid_alloc_cell = its header once `off` (start of its run) is known; the device kernel adds a warp's needs
// with one atomic, the host emulation cell by cell.
YRT_HD int pgrid_alloc_need(const PGridArrays& a, int n) { return (n > 0 && n <= a.d.max_list) ? n : 0; }
YRT_HD int2 pgrid_alloc_cell(const PGridArrays& a, int n, int off) {
    int2 c;
    c.x = 0; c.y = 0;
    if (n > a.d.max_list || (n > 0 && off + n > a.d.capacity)) c.y = -1;
    else if (n > 0) { c.x = off; c.y = n; }
    return c;
}
YRT_HD void pgrid_alloc_item(const PGridArrays& a, int cell) {
    const int n = a.cnt[cell];
    a.cnt[cell] = 0;
    const int need = pgrid_alloc_need(a, n);
    const int off = need ? YRT_ATOMIC_ADD(a.total, need) : 0;
    a.cells[cell] = pgrid_alloc_cell(a, n, off);
}

// chain nodes a list of n candidates needs at arity W: every node holds W - 1 candidates and the rest of the list, the last one up to W
YRT_HD int pgrid_chain_nodes(int n, int W) { return n <= W ? 1 : 1 + (n - W + W - 2) / (W - 1); }

YRT_HD nodebox pgrid_inst_nodebox(const PGridArrays& a, int slot) {
    const float4 bc = a.inst_box[2 * (size_t)slot], bh = a.inst_box[2 * (size_t)slot + 1];
    nodebox b;
    b.cx = bc.x; b.cy = bc.y; b.cz = bc.z; b.hx = bh.x; b.hy = bh.y; b.hz = bh.z;
    return b;
}

// One cell: order its candidates by (distance bound, slot), then write its list as a chain of W-wide nodes, last node
// first so that every node knows the box of "the rest" (the union of what follows it):
//   node k = [ candidate (W-1) k, ..., candidate (W-1) k + W - 2, node k + 1 ]      last node: up to W candidates
// Candidates are instance leaves (one instance each), exactly what a leaf of the instance tree looks like to the traversal.
template <int W>
YRT_HD void pgrid_emit_item(const PGridArrays& a, int cell) {
    const int2 c = a.cells[cell];
    if (c.y == 0) { a.roots[cell] = YRT_REF_DONE; return; }
    if (c.y < 0) { a.roots[cell] = pgrid_tlas_root(a); return; }
    unsigned long long* k = a.keys + c.x;
    const int n = c.y;
    for (int i = 1; i < n; i++) {
        const unsigned long long v = k[i];
        int j = i - 1;
        while (j >= 0 && k[j] > v) { k[j + 1] = k[j]; j--; }
        k[j + 1] = v;
    }
    const int nn = pgrid_chain_nodes(n, W);
    const int first = YRT_ATOMIC_ADD(a.node_total, nn);
    if (first + nn > a.node_capacity) { a.roots[cell] = pgrid_tlas_root(a); return; }
    const int stride = YRT_NODE_STRIDE(W);
    const int per_face = a.d.nx * a.d.ny, face = cell / per_face, in_face = cell - face * per_face;
    const PCellDirs cd = pgrid_cell_dirs(a.d, face, in_face % a.d.nx, in_face / a.d.nx);
    Box rest = box_invalid();
    int rest_ref = YRT_REF_DONE;
    for (int kn = nn - 1; kn >= 0; kn--) {
        const int c0 = (W - 1) * kn;                                   // first candidate of this node
        const int nc = kn == nn - 1 ? n - c0 : W - 1;                  // candidates it holds
        nodebox box[W];
        int ref[W];
        Box mine = box_invalid();
        for (int j = 0; j < nc; j++) {
            const int slot = (int)(unsigned)(k[c0 + j] & 0xffffffffull);
            ref[j] = make_leaf_ref(slot, 1);
            // the candidate's box, cut down to what a ray of this cell can reach of it (unless it surrounds the apex)
            const nodebox wb = pgrid_inst_nodebox(a, slot);
            const vec3 wlo = mk3(wb.cx - wb.hx, wb.cy - wb.hy, wb.cz - wb.hz), whi = mk3(wb.cx + wb.hx, wb.cy + wb.hy, wb.cz + wb.hz);
            vec3 clo = wlo - a.d.apex, chi = whi - a.d.apex;
            if (pgrid_box_distance(a.d, wlo, whi) >= a.d.near_all) pgrid_cell_clip(a.d, cd, clo, chi, clo, chi);
            clo = clo + a.d.apex; chi = chi + a.d.apex;
            nodebox cb;
            box_center_half(fmaxf(clo.x, wlo.x), fminf(chi.x, whi.x), cb.cx, cb.hx);
            box_center_half(fmaxf(clo.y, wlo.y), fminf(chi.y, whi.y), cb.cy, cb.hy);
            box_center_half(fmaxf(clo.z, wlo.z), fminf(chi.z, whi.z), cb.cz, cb.hz);
            box[j] = cb;
            box_expand(mine, mk3(cb.cx - cb.hx, cb.cy - cb.hy, cb.cz - cb.hz), mk3(cb.cx + cb.hx, cb.cy + cb.hy, cb.cz + cb.hz));
        }
        int nb = nc;
        if (kn < nn - 1) {
            nodebox rb;
            box_center_half(rest.lo.x, rest.hi.x, rb.cx, rb.hx);
            box_center_half(rest.lo.y, rest.hi.y, rb.cy, rb.hy);
            box_center_half(rest.lo.z, rest.hi.z, rb.cz, rb.hz);
            box[nb] = rb; ref[nb] = rest_ref; nb++;
        }
        const int idx = a.node_first + first + kn;
        node_pack<W>(a.nodes + (size_t)stride * idx, box, ref, nb);
        box_expand(rest, mine.lo, mine.hi);
        rest_ref = idx;
    }
    a.roots[cell] = rest_ref;
}
// List form of one cell: its candidates ordered by (distance bound, slot), each with the part of its box the cell's rays
// can reach; the traversal takes them in this order (Tracer<…, GRID>::next_candidate) and stops at the first whose bound
// lies beyond what the ray has already found.
YRT_HD void pgrid_emit_list_item(const PGridArrays& a, int cell) {
    const int2 c = a.cells[cell];
    if (c.y <= 0) return;
    unsigned long long* k = a.keys + c.x;
    const int n = c.y;
    for (int i = 1; i < n; i++) {
        const unsigned long long v = k[i];
        int j = i - 1;
        while (j >= 0 && k[j] > v) { k[j + 1] = k[j]; j--; }
        k[j + 1] = v;
    }
    const int per_face = a.d.nx * a.d.ny, face = cell / per_face, in_face = cell - face * per_face;
    const PCellDirs cd = pgrid_cell_dirs(a.d, face, in_face % a.d.nx, in_face / a.d.nx);
    for (int i = 0; i < n; i++) {
        const int slot = (int)(unsigned)(k[i] & 0xffffffffull);
        const nodebox wb = pgrid_inst_nodebox(a, slot);
        const vec3 wlo = mk3(wb.cx - wb.hx, wb.cy - wb.hy, wb.cz - wb.hz), whi = mk3(wb.cx + wb.hx, wb.cy + wb.hy, wb.cz + wb.hz);
        vec3 clo = wlo - a.d.apex, chi = whi - a.d.apex;
        if (pgrid_box_distance(a.d, wlo, whi) >= a.d.near_all) pgrid_cell_clip(a.d, cd, clo, chi, clo, chi);
        clo = clo + a.d.apex; chi = chi + a.d.apex;
        nodebox cb;
        box_center_half(fmaxf(clo.x, wlo.x), fminf(chi.x, whi.x), cb.cx, cb.hx);
        box_center_half(fmaxf(clo.y, wlo.y), fminf(chi.y, whi.y), cb.cy, cb.hy);
        box_center_half(fmaxf(clo.z, wlo.z), fminf(chi.z, whi.z), cb.cz, cb.hz);
        float4* e = a.entries + 2 * ((size_t)c.x + i);
        e[0] = mk4(cb.cx, cb.cy, cb.cz, int_as_float(slot));
        e[1] = mk4(cb.hx, cb.hy, cb.hz, int_as_float((int)(unsigned)(k[i] >> 32)));
    }
}
// List form with fixed_k slots per cell, one item per (cell, slot): the header of the cell (slot 0), and — if the cell is
// served and the slot holds a candidate — that candidate's entry at its place in the order, i.e. at the number of the
// cell's keys below its own (the keys of a cell differ at least in the instance slot).  No per-cell sort, no serial loop
// over a cell's candidates: one short thread per entry.
YRT_HD void pgrid_emit_slot_item(const PGridArrays& a, int t) {
    const int K = a.fixed_k, cell = t / K, i = t - cell * K;
    const int n = a.cnt[cell];
    const bool served = n <= K && n <= a.d.max_list;
    if (i == 0) { int2 c; c.x = cell * K; c.y = served ? n : -1; a.cells[cell] = c; }
    if (!served || i >= n) return;
    const unsigned long long* k = a.keys + (size_t)cell * K;
    const unsigned long long mine = k[i];
    int rank = 0;
    for (int j = 0; j < n; j++) rank += k[j] < mine;
    const int slot = (int)(unsigned)(mine & 0xffffffffull);
    const int per_face = a.d.nx * a.d.ny, face = cell / per_face, in_face = cell - face * per_face;
    const PCellDirs cd = pgrid_cell_dirs(a.d, face, in_face % a.d.nx, in_face / a.d.nx);
    const nodebox wb = pgrid_inst_nodebox(a, slot);
    const vec3 wlo = mk3(wb.cx - wb.hx, wb.cy - wb.hy, wb.cz - wb.hz), whi = mk3(wb.cx + wb.hx, wb.cy + wb.hy, wb.cz + wb.hz);
    vec3 clo = wlo - a.d.apex, chi = whi - a.d.apex;
    if (pgrid_box_distance(a.d, wlo, whi) >= a.d.near_all) pgrid_cell_clip(a.d, cd, clo, chi, clo, chi);
    clo = clo + a.d.apex; chi = chi + a.d.apex;
    nodebox cb;
    box_center_half(fmaxf(clo.x, wlo.x), fminf(chi.x, whi.x), cb.cx, cb.hx);
    box_center_half(fmaxf(clo.y, wlo.y), fminf(chi.y, whi.y), cb.cy, cb.hy);
    box_center_half(fmaxf(clo.z, wlo.z), fminf(chi.z, whi.z), cb.cz, cb.hz);
    float4* e = a.entries + 2 * ((size_t)cell * K + rank);
    e[0] = mk4(cb.cx, cb.cy, cb.cz, int_as_float(slot));
    e[1] = mk4(cb.hx, cb.hy, cb.hz, int_as_float((int)(unsigned)(mine >> 32)));
}

YRT_HD void pgrid_emit_item(const PGridArrays& a, int cell) {
    if (a.as_list) pgrid_emit_list_item(a, cell);
    else if (a.arity == 4) pgrid_emit_item<4>(a, cell);
    else pgrid_emit_item<2>(a, cell);
}

// ---- descriptions (host) ----------------------------------------------------------------------------------------------
// u = 2^-24.  `extent` bounds |coordinate| of everything the rays can touch (scene box and apex).
YRT_HD float pgrid_abs_error(float extent) { return 16.0f * 5.9604645e-8f * extent; }   // 16 u extent: the sums (X - apex) and the normalisation round <= 8 times at magnitudes <= extent

// camera grid for a width x height frame; mode 0 when the camera frame is not (nearly) orthonormal
YRT_HD PGridDesc pgrid_camera_desc(const camera_k& cam, int width, int height, int shift, float extent, int capacity) {
    PGridDesc d;
    d.mode = 1;
    const int cell = 1 << shift;
    d.nx = (width + cell - 1) >> shift; d.ny = (height + cell - 1) >> shift;
    d.n_cells = d.nx * d.ny;
    d.apex = cam.frame.o;
    d.ex = cam.frame.x; d.ey = cam.frame.y * -1.0f; d.ez = cam.frame.z * -1.0f;
    d.sx = (float)width * cam.focus / cam.w; d.sy = (float)height * cam.focus / cam.h;
    d.cx = 0.5f * (float)width; d.cy = 0.5f * (float)height;
    d.inv_cell = 1.0f / (float)cell;
    // deviation of the camera frame from an orthonormal one
    float dev = 0.f;
    const vec3 e[3] = {d.ex, d.ey, d.ez};
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) dev = fmaxf(dev, fabsf(dot(e[i], e[j]) - (i == j ? 1.0f : 0.0f)));
    const float err = pgrid_abs_error(extent + fabsf(cam.w) + fabsf(cam.h) + fabsf(cam.focus));
    // error of q - o (eval_camera) measured in pixels of the focal plane, + the frame's deviation over the whole image
    const float eps_px = err / fminf(fabsf(cam.w) / (float)width, fabsf(cam.h) / (float)height);
    d.margin = 0.5f + 2.0f * eps_px + 8.0f * dev * (float)(width + height);
    d.near_all = fmaxf(1.0e-3f * fabsf(cam.focus), 64.0f * err);
    d.skip_within = 0.f;
    // (instance frames count as rigid up to 1e-4 in their squared lengths: a local t may differ from the distance by 5e-5 of it)
    d.dist_scale = 1.0f - 1.0e-3f; d.dist_bias = 4.0f * err;
    d.max_list = YRT_PGRID_MAX_LIST; d.capacity = capacity;
    if (!(dev <= 1.0e-4f) || !(d.margin <= 4.0f) || !(cam.focus > 0.f) || !(cam.w > 0.f) || !(cam.h > 0.f)) d.mode = 0;
    return d;
}

// cube grid around a point light at `apex`
YRT_HD PGridDesc pgrid_cube_desc(const vec3& apex, int R, float extent, int capacity) {
    PGridDesc d;
    d.mode = 2;
    d.nx = d.ny = R; d.n_cells = 6 * R * R;
    d.apex = apex;
    d.ex = mk3(1, 0, 0); d.ey = mk3(0, 1, 0); d.ez = mk3(0, 0, 1);
    d.sx = d.sy = d.cx = d.cy = 0.f; d.inv_cell = 0.f;
    const float err = pgrid_abs_error(extent);
    // a point at distance s from the light is seen from it under a direction that differs from the ray's by <= err / s:
    // boxes nearer than near_all go everywhere, the others get 2 err / near_all of margin (|du/dangle| <= 2 on a face)
    d.margin = fminf(0.25f * 2.0f / (float)R, 2.0e-3f);
    d.near_all = fmaxf(2.0f * err / (0.5f * d.margin), 0.02f);
    d.margin += 1.0e-5f;
    d.skip_within = 0.01f * (1.0f - 1.0e-3f) - 4.0f * err;
    d.dist_scale = 1.0f - 1.0e-3f; d.dist_bias = 4.0f * err;
    d.max_list = YRT_PGRID_MAX_LIST; d.capacity = capacity;
    return d;
}

}  // namespace yrt
