// host_emu.cu — TEST INFRASTRUCTURE, never part of the product path.
//
// Compiles the __host__ __device__ code of yocto_raytracing_b200/csrc (math, LBVH item functions,
// traversal, shading) for the CPU and drives it with plain loops, so that the device logic can be
// checked against the reference where there is no GPU: same Morton keys / Karras topology / refit
// / node emit as the CUDA build (std::stable_sort stands in for the device radix sort), same
// trace_ray / shade_hit code as the kernels.  libyrt_b200.so never links or calls this.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include "yrt_internal.h"
#include "yrt_shade.cuh"
#include "yrt_trace.cuh"

using namespace yrt;

namespace {

struct EmuLbvh {
    std::vector<float4> nodes2, nodes4;
    std::vector<int> order, seg_root, seg_depth, seg_need2, seg_need4, seg_box_lo, seg_box_hi;
};

void emu_lbvh(int n, int n_seg, std::vector<float4>& lo, std::vector<float4>& hi, const int* seg_of, const int* seg_first, int leaf_size,
              EmuLbvh& out, int ref_offset, int size_bits, int rotate_rounds, int rotate_pairs) {
    size_t ni = n > 1 ? n - 1 : 1;
    std::vector<int> cent_lo(3 * n_seg), cent_hi(3 * n_seg);
    out.seg_box_lo.assign(3 * n_seg, 0);
    out.seg_box_hi.assign(3 * n_seg, 0);
    std::vector<unsigned long long> keys(std::max(n, 1));
    out.order.assign(std::max(n, 1), 0);
    std::vector<int> left(ni), right(ni), rfirst(ni), rlast(ni), pint(ni), pleaf(std::max(n, 1), -1), flags(ni);
    std::vector<float4> nlo(ni), nhi(ni);
    std::vector<int> count(ni), new_slot(std::max(n, 1)), order_tmp(std::max(n, 1)), pleaf_tmp(std::max(n, 1));
    out.nodes2.assign(YRT_NODE_STRIDE(2) * ni, mk4(0, 0, 0, 0));
    out.nodes4.assign(YRT_NODE_STRIDE(4) * ni, mk4(0, 0, 0, 0));
    out.seg_root.assign(n_seg, 0);
    out.seg_depth.assign(n_seg, 0);
    out.seg_need2.assign(n_seg, 0);
    out.seg_need4.assign(n_seg, 0);
    std::vector<int> need2(ni, 0), need4(ni, 0);
    LbvhArrays a;
    a.n = n; a.n_seg = n_seg; a.box_lo = lo.data(); a.box_hi = hi.data(); a.seg_of = seg_of; a.seg_first = seg_first;
    a.seg_cent_lo = cent_lo.data(); a.seg_cent_hi = cent_hi.data(); a.seg_box_lo = out.seg_box_lo.data(); a.seg_box_hi = out.seg_box_hi.data();
    a.keys = keys.data(); a.order = out.order.data(); a.left = left.data(); a.right = right.data();
    a.range_first = rfirst.data(); a.range_last = rlast.data(); a.parent_int = pint.data(); a.parent_leaf = pleaf.data();
    a.count = count.data(); a.new_slot = new_slot.data(); a.order_tmp = order_tmp.data(); a.parent_leaf_tmp = pleaf_tmp.data();
    a.flags = flags.data(); a.node_lo = nlo.data(); a.node_hi = nhi.data(); a.nodes2 = out.nodes2.data(); a.nodes4 = out.nodes4.data();
    a.seg_root = out.seg_root.data(); a.seg_depth = out.seg_depth.data(); a.seg_need2 = out.seg_need2.data(); a.seg_need4 = out.seg_need4.data(); a.need2 = need2.data(); a.need4 = need4.data(); a.leaf_size = leaf_size; a.ref_offset = ref_offset; a.size_bits = size_bits;
    a.rotate_pairs = rotate_pairs;
    for (int s = 0; s < n_seg; s++) seg_bounds_init_item(a, s);
    for (int i = 0; i < n; i++) seg_bounds_item(a, i);
    for (int i = 0; i < n; i++) morton_item(a, i);
    {   // stands in for the stable LSD radix sort of (key, order) pairs
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return keys[x] < keys[y]; });
        std::vector<unsigned long long> k2(n);
        std::vector<int> o2(n);
        for (int k = 0; k < n; k++) { k2[k] = keys[idx[k]]; o2[k] = out.order[idx[k]]; }
        std::copy(k2.begin(), k2.end(), keys.begin());
        std::copy(o2.begin(), o2.end(), out.order.begin());
    }
    if (n > 1) {
        for (int i = 0; i < n - 1; i++) karras_item(a, i);
        for (int i = 0; i < n; i++) refit_item(a, i);
        for (int r = 0; r < rotate_rounds; r++) {
            std::fill(flags.begin(), flags.end(), 0);
            for (int i = 0; i < n; i++) rotate_refit_item(a, i);
        }
        if (rotate_rounds > 0) {
            for (int i = 0; i < n; i++) relayout_slot_item(a, i);
            for (int i = 0; i < n; i++) relayout_move_item(a, i);
            for (int i = 0; i < n - 1; i++) relayout_refs_item(a, i);
            for (int i = 0; i < n; i++) relayout_copy_item(a, i);
        }
        for (int i = 0; i < n - 1; i++) emit_item(a, i);
    }
    for (int s = 0; s < n_seg; s++) single_root_item(a, s);
    if (n > 1) {
        for (int i = 0; i < n; i++) depth_item(a, i);
        std::fill(flags.begin(), flags.end(), 0);
        for (int i = 0; i < n; i++) stackneed_item(a, i);
    }
}

// one apex grid built by serial loops over the item functions of yrt_pgrid.cuh (the device runs the same functions per thread)
struct EmuGrid {
    PGridDesc d;
    std::vector<int> roots;           // chain form
    std::vector<int2> cells;          // list form
    std::vector<float4> entries;
    long long n_entries = 0, n_fallback_cells = 0, n_nodes = 0;
    GridRef ref() const {
        GridRef g = gridref_none();
        if (d.mode != 0) { g.roots = roots.data(); g.cells = cells.data(); g.entries = entries.data(); g.nx = d.nx; g.shift = 0; }
        return g;
    }
};

// chain form (nodes != null): chain nodes go to records [node_first, node_first + node_cap) of `nodes` (arity W); *node_total counts
// them (shared by the lights).  List form (nodes == null): out.cells + out.entries.
void emu_build_grid(const PGridDesc& d, const std::vector<float4>& inst_box, int n_inst, int arity, float4* nodes, int node_first, int node_cap,
                    int* node_total, int tlas_root, EmuGrid& out) {
    out.d = d;
    out.roots.assign(std::max(d.n_cells, 1), tlas_root);
    out.cells.assign(std::max(d.n_cells, 1), int2{0, 0});
    out.entries.assign(nodes ? 2 : 2 * (size_t)std::max(d.capacity, 1), mk4(0, 0, 0, 0));
    out.n_entries = out.n_fallback_cells = out.n_nodes = 0;
    if (d.mode == 0) return;
    std::vector<int> cnt((size_t)d.n_cells + 2 + n_inst, 0);
    std::vector<int2>& cells = out.cells;
    std::vector<unsigned long long> keys(std::max(d.capacity, 1));
    int dummy_total = 0;
    if (!node_total) node_total = &dummy_total;
    PGridArrays a;
    a.d = d; a.inst_box = inst_box.data(); a.n_inst = n_inst;
    a.cnt = cnt.data(); a.total = a.cnt + d.n_cells; a.big = a.cnt + d.n_cells + 1;
    a.cells = cells.data(); a.keys = keys.data(); a.roots = out.roots.data();
    a.nodes = nodes; a.arity = arity; a.node_first = node_first; a.node_capacity = node_cap; a.node_total = node_total;
    a.as_list = nodes ? 0 : 1; a.entries = out.entries.data();
    a.fixed_k = nodes ? 0 : 16;
    if (a.fixed_k) {   // (mirrors camera_grid_enqueue: one scatter pass into the cells' own slots, one item per slot)
        a.d.max_list = std::min(a.d.max_list, a.fixed_k);
        keys.assign((size_t)d.n_cells * a.fixed_k, 0ull); a.keys = keys.data();
        out.entries.assign(2 * (size_t)d.n_cells * a.fixed_k, mk4(0, 0, 0, 0)); a.entries = out.entries.data();
        for (int k = 0; k < n_inst; k++) pgrid_scatter_item(a, k, 0, 1, false, false);
        for (int b = 0; b < a.big[0]; b++) pgrid_scatter_item(a, a.big[1 + b], 0, 1, false, true);
        for (int t = 0; t < d.n_cells * a.fixed_k; t++) pgrid_emit_slot_item(a, t);
        for (int c = 0; c < d.n_cells; c++) { out.n_fallback_cells += cells[c].y < 0; out.n_entries += std::max(cells[c].y, 0); }
        return;
    }
    a.tlas_root = tlas_root; a.tlas_root_dev = nullptr;
    const int before = *node_total;
    for (int fill = 0; fill < 2; fill++) {
        for (int k = 0; k < n_inst; k++) pgrid_scatter_item(a, k, 0, 1, fill != 0, false);
        for (int b = 0; b < a.big[0]; b++) pgrid_scatter_item(a, a.big[1 + b], 0, 1, fill != 0, true);
        if (!fill) for (int c = 0; c < d.n_cells; c++) pgrid_alloc_item(a, c);
    }
    for (int c = 0; c < d.n_cells; c++) pgrid_emit_item(a, c);
    out.n_entries = *a.total;
    out.n_nodes = *node_total - before;
    for (int c = 0; c < d.n_cells; c++) out.n_fallback_cells += cells[c].y < 0 || (nodes && cells[c].y > 0 && out.roots[c] == tlas_root);
}

struct EmuScene {
    HostScene hs;
    std::vector<float4> inst_box;
    std::vector<EmuGrid> light_grids;     // one per light (mode 0: none)
    LightGrids lg;
    int cam_shift = -1;                   // >= 0: trace_primary / render build a camera grid per call
    float extent = 0.f;
    int tree_nodes = 0;                   // records of the two trees in either node array; the grids' chain nodes follow
    int light_node_cap = 0;
    EmuLbvh blas, tlas;
    std::vector<float4> prim_recs, prim_attrs, inst_recs;
    std::vector<int> prim_rank, inst_rank;
    std::vector<float4> nodes2, nodes4;   // BLAS nodes then TLAS nodes, like the device arrays
    SceneView view;
    RefTlas ref = {nullptr, nullptr, nullptr, 0};   // scenes with non-rigid instance frames (mirrors build_device_scene + k_slot_of_inst)
    std::vector<int> slot_of_inst;
    int blas_depth = 0, tlas_depth = 0, stack_need = 0;
};

int emu_build(const yrt_scene_desc* d, EmuScene& es, int leaf_blas, int leaf_tlas) {
    YRT_TRY(host_scene_from_desc(d, es.hs));
    HostScene& hs = es.hs;
    GeomView g;
    g.shape_kind = hs.shape_kind.data(); g.shape_elem_off = hs.shape_elem_off.data(); g.shape_elem_cnt = hs.shape_elem_cnt.data();
    g.shape_vert_off = hs.shape_vert_off.data(); g.shape_prim_off = hs.shape_prim_off.data(); g.elem_idx = hs.elem_idx.data();
    g.pos = hs.pos.data(); g.norm = hs.norm.data(); g.uv = hs.uv.data(); g.radius = hs.radius.data();
    g.prim_shape = hs.prim_shape.data(); g.n_prims = hs.n_prims;
    int np = hs.n_prims;
    std::vector<float4> plo(std::max(np, 1)), phi(std::max(np, 1));
    for (int i = 0; i < np; i++) {
        Box b = prim_bounds(g, i);
        plo[i] = mk4(b.lo.x, b.lo.y, b.lo.z, 0.f);
        phi[i] = mk4(b.hi.x, b.hi.y, b.hi.z, 0.f);
    }
    int nseg = std::max(hs.n_shapes, 1);
    emu_lbvh(np, nseg, plo, phi, g.prim_shape, g.shape_prim_off, leaf_blas, es.blas, 0, getenv("YRT_SIZE_BITS_BLAS") ? atoi(getenv("YRT_SIZE_BITS_BLAS")) : YRT_SIZE_BITS_BLAS,
             getenv("YRT_ROTATE_BLAS") ? atoi(getenv("YRT_ROTATE_BLAS")) : YRT_ROTATE_ROUNDS_BLAS,
             getenv("YRT_ROTATE_PAIRS_BLAS") ? atoi(getenv("YRT_ROTATE_PAIRS_BLAS")) : YRT_ROTATE_PAIRS_BLAS);
    // prim + attribute records in BLAS leaf order (mirrors k_gather_prims)
    es.prim_recs.assign(3 * (size_t)std::max(np, 1), mk4(0, 0, 0, 0));
    es.prim_attrs.assign(YRT_ATTR_STRIDE * (size_t)std::max(np, 1), mk4(0, 0, 0, 0));
    es.prim_rank.assign(std::max(np, 1), 0);
    for (int k = 0; k < np; k++) {
        int gp = es.blas.order[k];
        es.prim_rank[k] = hs.prim_rank[gp];
        int s = g.prim_shape[gp], e = gp - g.shape_prim_off[s], kind = g.shape_kind[s], vo = g.shape_vert_off[s];
        int nv = kind == 0 ? 3 : (kind == 1 ? 2 : 1);
        const int* t = g.elem_idx + g.shape_elem_off[s] + (size_t)nv * e;
        int v[3] = {vo + t[0], vo + t[nv > 1 ? 1 : 0], vo + t[nv > 2 ? 2 : 0]};
        vec3 p0 = ld3(g.pos, v[0]), p1 = ld3(g.pos, v[1]), p2 = ld3(g.pos, v[2]);
        vec3 n0 = ld3(g.norm, v[0]), n1 = ld3(g.norm, v[1]), n2 = ld3(g.norm, v[2]);
        bool huv = hs.shape_has_uv[s] != 0;
        float u0 = huv ? g.uv[2 * v[0]] : 0.f, w0 = huv ? g.uv[2 * v[0] + 1] : 0.f;
        float u1 = huv ? g.uv[2 * v[1]] : 0.f, w1 = huv ? g.uv[2 * v[1] + 1] : 0.f;
        float u2 = huv ? g.uv[2 * v[2]] : 0.f, w2 = huv ? g.uv[2 * v[2] + 1] : 0.f;
        float4* pr = &es.prim_recs[3 * (size_t)k];
        float4* ar = &es.prim_attrs[YRT_ATTR_STRIDE * (size_t)k];
        pack_prim(kind, e, p0, p1, p2, kind == 0 ? 0.f : g.radius[v[0]], kind == 1 ? g.radius[v[1]] : 0.f, pr, ar);
        ar[0] = mk4(n0.x, n0.y, n0.z, u0); ar[1] = mk4(n1.x, n1.y, n1.z, w0); ar[2] = mk4(n2.x, n2.y, n2.z, u1); ar[3] = mk4(w1, u2, w2, 0.f);
    }
    // TLAS (mirrors k_inst_boxes / k_inst_recs)
    int na0 = (int)hs.active_inst.size();
    struct Item { int active; int root; Box box; };
    std::vector<Item> items;
    for (int a = 0; a < na0; a++) {
        int inst = hs.active_inst[a], s = hs.inst_shape[inst];
        Box b;
        b.lo = mk3(ordered_to_float(es.blas.seg_box_lo[3 * s]), ordered_to_float(es.blas.seg_box_lo[3 * s + 1]), ordered_to_float(es.blas.seg_box_lo[3 * s + 2]));
        b.hi = mk3(ordered_to_float(es.blas.seg_box_hi[3 * s]), ordered_to_float(es.blas.seg_box_hi[3 * s + 1]), ordered_to_float(es.blas.seg_box_hi[3 * s + 2]));
        items.push_back({a, es.blas.seg_root[s], b});
    }
    int na = (int)items.size();
    std::vector<float4> ilo(std::max(na, 1)), ihi(std::max(na, 1));
    for (int a = 0; a < na; a++) {
        int inst = hs.active_inst[items[a].active];
        const float* fr = &hs.inst_frame[12 * (size_t)inst];
        frame3 f;
        f.x = mk3(fr[0], fr[1], fr[2]); f.y = mk3(fr[3], fr[4], fr[5]); f.z = mk3(fr[6], fr[7], fr[8]); f.o = mk3(fr[9], fr[10], fr[11]);
        Box w = instance_bounds(f, items[a].box);
        ilo[a] = mk4(w.lo.x, w.lo.y, w.lo.z, 0.f);
        ihi[a] = mk4(w.hi.x, w.hi.y, w.hi.z, 0.f);
    }
    std::vector<int> seg_of(std::max(na, 1), 0);
    int sf[2] = {0, na};
    int nb_int = np > 1 ? np - 1 : 0;
    emu_lbvh(na, 1, ilo, ihi, seg_of.data(), sf, leaf_tlas, es.tlas, nb_int, getenv("YRT_SIZE_BITS_TLAS") ? atoi(getenv("YRT_SIZE_BITS_TLAS")) : YRT_SIZE_BITS_TLAS,
             getenv("YRT_ROTATE_TLAS") ? atoi(getenv("YRT_ROTATE_TLAS")) : YRT_ROTATE_ROUNDS_TLAS,
             getenv("YRT_ROTATE_PAIRS_TLAS") ? atoi(getenv("YRT_ROTATE_PAIRS_TLAS")) : YRT_ROTATE_PAIRS_TLAS);
    es.nodes2.assign(es.blas.nodes2.begin(), es.blas.nodes2.begin() + YRT_NODE_STRIDE(2) * (size_t)nb_int);
    es.nodes2.insert(es.nodes2.end(), es.tlas.nodes2.begin(), es.tlas.nodes2.end());
    es.nodes4.assign(es.blas.nodes4.begin(), es.blas.nodes4.begin() + YRT_NODE_STRIDE(4) * (size_t)nb_int);
    es.nodes4.insert(es.nodes4.end(), es.tlas.nodes4.begin(), es.tlas.nodes4.end());
    es.inst_recs.assign(4 * (size_t)std::max(na, 1), mk4(0, 0, 0, 0));
    es.inst_rank.assign(std::max(na, 1), 0);
    for (int k = 0; k < na; k++) {
        const Item& it = items[es.tlas.order[k]];
        int inst = hs.active_inst[it.active], s = hs.inst_shape[inst];
        es.inst_rank[k] = hs.inst_rank[inst];
        const float* fr = &hs.inst_frame[12 * (size_t)inst];
        float4* r = &es.inst_recs[4 * (size_t)k];
        r[0] = mk4(fr[0], fr[1], fr[2], int_as_float(it.root));
        r[1] = mk4(fr[3], fr[4], fr[5], int_as_float(inst));
        r[2] = mk4(fr[6], fr[7], fr[8], int_as_float(hs.inst_mat[inst]));
        r[3] = mk4(fr[9], fr[10], fr[11], int_as_float((int)((unsigned)s | ((unsigned)hs.shape_kind[s] << 28))));
    }
    es.inst_box.assign(2 * (size_t)std::max(na, 1), mk4(0, 0, 0, 0));
    for (int k = 0; k < na; k++) {   // mirrors k_inst_box
        const float4 l = ilo[es.tlas.order[k]], h = ihi[es.tlas.order[k]];
        float cx, cy, cz, hx, hy, hz;
        box_center_half(l.x, h.x, cx, hx); box_center_half(l.y, h.y, cy, hy); box_center_half(l.z, h.z, cz, hz);
        es.inst_box[2 * (size_t)k] = mk4(cx, cy, cz, 0.f);
        es.inst_box[2 * (size_t)k + 1] = mk4(hx, hy, hz, 0.f);
    }
    if (!hs.all_rigid && hs.n_instances > 0 && !getenv("YRT_EMU_NO_REF_TLAS")) {   // (the switch: tests show that the LBVH alone does not reproduce such scenes)
        es.slot_of_inst.assign(hs.n_instances, -1);
        for (int k = 0; k < na; k++) es.slot_of_inst[hs.active_inst[items[es.tlas.order[k]].active]] = k;
        es.ref.nodes = hs.ref_nodes.data(); es.ref.leaf_inst = hs.ref_leaf_inst.data(); es.ref.slot_of_inst = es.slot_of_inst.data();
        es.ref.n_nodes = (int)(hs.ref_nodes.size() / 2);
    }
    es.extent = hs.extent;
    es.tree_nodes = nb_int + (na > 1 ? na - 1 : 1) + 2;
    for (int k = 0; k < YRT_MAX_LIGHT_GRIDS; k++) es.lg.g[k] = gridref_none();
    es.blas_depth = es.blas.seg_depth.empty() ? 0 : *std::max_element(es.blas.seg_depth.begin(), es.blas.seg_depth.end());
    es.tlas_depth = es.tlas.seg_depth[0];
    {
        const int n2 = es.tlas.seg_need2[0] + (es.blas.seg_need2.empty() ? 0 : *std::max_element(es.blas.seg_need2.begin(), es.blas.seg_need2.end())) + 3;
        const int n4 = es.tlas.seg_need4[0] + (es.blas.seg_need4.empty() ? 0 : *std::max_element(es.blas.seg_need4.begin(), es.blas.seg_need4.end())) + 3;
        es.stack_need = std::max(YRT_WIDE_CLOSEST == 4 ? n4 : n2, YRT_WIDE_ANY == 4 ? n4 : n2);
    }
    if (es.stack_need > YRT_STACK_CAP) { set_error("emu: tree too deep"); return YRT_ERR_UNSUPPORTED; }
    SceneView& v = es.view;
    v.nodes2 = es.nodes2.data(); v.nodes4 = es.nodes4.data(); v.inst_recs = es.inst_recs.data(); v.inst_box = es.inst_box.data();
    v.prim_recs = es.prim_recs.data(); v.prim_attrs = es.prim_attrs.data(); v.mat_recs = hs.mat_recs.data();
    v.light_recs = hs.light_recs.data(); v.tex_rgba8 = hs.tex_rgba8.data(); v.tex_info = hs.tex_info.data();
    v.inst_rank = es.inst_rank.data(); v.prim_rank = es.prim_rank.data();
    v.srgb_lut = hs.srgb_lut; v.tlas_root = es.tlas.seg_root[0]; v.n_lights = (int)hs.light_inst.size(); v.n_active_instances = na;
    return YRT_OK;
}

void hit_to_ids(const SceneView& sv, const HitRec& h, int* ids) {
    if (h.si < 0) { ids[0] = ids[1] = ids[2] = -1; return; }
    const float4* ir = sv.inst_recs + 4 * (size_t)h.si;
    ids[0] = float_as_int(ir[1].w);
    ids[1] = float_as_int(ir[3].w) & 0x0fffffff;
    ids[2] = float_as_int(sv.prim_recs[3 * (size_t)h.prim].w);
}

camera_k emu_camera(const yrt_camera* cam) {
    camera_k ck;
    const float* f = cam->frame;
    ck.frame.x = mk3(f[0], f[1], f[2]); ck.frame.y = mk3(f[3], f[4], f[5]); ck.frame.z = mk3(f[6], f[7], f[8]); ck.frame.o = mk3(f[9], f[10], f[11]);
    ck.h = 2.0f * cam->focus * tanf(cam->fovy / 2.0f); ck.w = ck.h * cam->aspect; ck.focus = cam->focus;
    return ck;
}

// the camera grid of one call (mirrors camera_grid_enqueue in yrt_build.cu)
void emu_camera_grid(EmuScene& es, const camera_k& ck, int width, int height, EmuGrid& g) {
    g.d.mode = 0;
    if (es.cam_shift < 0 || !es.hs.all_rigid || es.view.n_active_instances <= 0) return;
    const float reach = es.extent + fmaxf(fmaxf(fabsf(ck.frame.o.x), fabsf(ck.frame.o.y)), fabsf(ck.frame.o.z));
    const long long cells = (long long)((width + (1 << es.cam_shift) - 1) >> es.cam_shift) * ((height + (1 << es.cam_shift) - 1) >> es.cam_shift);
    PGridDesc d = pgrid_camera_desc(ck, width, height, es.cam_shift, reach, (int)(16 * cells + 65536));
    emu_build_grid(d, es.inst_box, es.view.n_active_instances, 0, nullptr, 0, 0, nullptr, es.view.tlas_root, g);
}
GridRef emu_camera_ref(const EmuScene& es, const EmuGrid& g) {
    GridRef r = g.ref();
    r.shift = es.cam_shift;
    return r;
}

// the kernels' choice (run_batch / intersect_rays_device in yrt_render.cu): scenes with non-rigid frames walk the reference's instance tree
bool emu_closest(const EmuScene& es, const GridRef& cg, const ray3& ray, int i, int j, bool primary, HitRec& h, int* stack, TraceCounters* tc) {
    if (es.ref.n_nodes > 0) { int tstack[YRT_REF_TLAS_STACK]; return trace_ray_ref<false>(es.view, es.ref, ray, h, stack, tstack, tc); }
    if (primary) { trace_camera_ray(es.view, cg, ray, i, j, h, stack, tc); return h.si >= 0; }
    return trace_ray<false>(es.view, ray, h, stack, tc);
}
bool emu_any(const EmuScene& es, const ray3& ray, HitRec& h, int* stack, TraceCounters* tc) {
    if (es.ref.n_nodes > 0) { int tstack[YRT_REF_TLAS_STACK]; return trace_ray_ref<true>(es.view, es.ref, ray, h, stack, tstack, tc); }
    return trace_ray<true>(es.view, ray, h, stack, tc);
}

}  // namespace

extern "C" {

const char* emu_last_error(void) { return get_error(); }

// apex grids on / off: light_R > 0 builds one cube grid per point light whose frame does not rotate (like build_device_scene),
// cam_shift >= 0 makes trace_primary / render build a camera grid per call (cells of 2^shift pixels).  stats (optional, 4
// int64): light-grid entries, light-grid cells that fall back to the tree, lights with a grid, 0
int emu_set_grids(void* p, int light_R, int cam_shift, int64_t* stats) {
    EmuScene* es = (EmuScene*)p;
    es->cam_shift = cam_shift;
    es->light_grids.clear();
    long long ne = 0, nf = 0, nl = 0, nn = 0;
    for (int k = 0; k < YRT_MAX_LIGHT_GRIDS; k++) es->lg.g[k] = gridref_none();
    const int n_lights = es->view.n_lights, na = es->view.n_active_instances;
    const bool lights_on = light_R > 0 && es->hs.all_rigid && na > 0;
    // room for the chain nodes behind the trees (same regions as build_device_scene)
    const int n_grids = lights_on ? std::min(n_lights, YRT_MAX_LIGHT_GRIDS) : 0;
    es->light_node_cap = n_grids ? (getenv("YRT_LIGHT_GRID_NODES") ? atoi(getenv("YRT_LIGHT_GRID_NODES")) : n_grids * 6 * light_R * light_R * 3 / 4 + 1024) : 0;
    const size_t total_nodes = (size_t)es->tree_nodes + es->light_node_cap;
    es->nodes2.resize(YRT_NODE_STRIDE(2) * total_nodes, mk4(0, 0, 0, 0));
    es->nodes4.resize(YRT_NODE_STRIDE(4) * total_nodes, mk4(0, 0, 0, 0));
    es->view.nodes2 = es->nodes2.data(); es->view.nodes4 = es->nodes4.data();
    if (lights_on) {
        es->light_grids.resize(n_grids);
        int node_total = 0;
        std::vector<float4>& arr = YRT_WIDE_ANY == 4 ? es->nodes4 : es->nodes2;
        for (int k = 0; k < n_grids; k++) {
            const float4* lr = &es->hs.light_recs[5 * (size_t)k];
            const bool identity = lr[0].x == 1.f && lr[0].y == 0.f && lr[0].z == 0.f && lr[1].x == 0.f && lr[1].y == 1.f && lr[1].z == 0.f &&
                                  lr[2].x == 0.f && lr[2].y == 0.f && lr[2].z == 1.f;
            es->light_grids[k].d.mode = 0;
            if (!identity) continue;
            const vec3 apex = mk3(lr[4].x + lr[3].x, lr[4].y + lr[3].y, lr[4].z + lr[3].z);
            const float reach = es->extent + fmaxf(fmaxf(fabsf(apex.x), fabsf(apex.y)), fabsf(apex.z));
            emu_build_grid(pgrid_cube_desc(apex, light_R, reach, getenv("YRT_LIGHT_GRID_KEYS") ? atoi(getenv("YRT_LIGHT_GRID_KEYS")) : 4 * 6 * light_R * light_R), es->inst_box, na, YRT_WIDE_ANY, arr.data(), es->tree_nodes,
                           es->light_node_cap, &node_total, es->view.tlas_root, es->light_grids[k]);
            es->lg.g[k] = es->light_grids[k].ref();
            ne += es->light_grids[k].n_entries; nf += es->light_grids[k].n_fallback_cells; nl++; nn += es->light_grids[k].n_nodes;
        }
    }
    if (stats) { stats[0] = ne; stats[1] = nf; stats[2] = nl; stats[3] = nn; }
    return YRT_OK;
}

int emu_scene_create(const yrt_scene_desc* d, int leaf_blas, int leaf_tlas, void** out) {
    EmuScene* es = new EmuScene();
    int st = emu_build(d, *es, leaf_blas > 0 ? leaf_blas : YRT_LEAF_SIZE_BLAS, leaf_tlas > 0 ? leaf_tlas : YRT_LEAF_SIZE_TLAS);
    if (st != YRT_OK) { delete es; return st; }
    *out = es;
    return YRT_OK;
}
void emu_scene_destroy(void* p) { delete (EmuScene*)p; }
int emu_scene_info(void* p, int64_t out[8]) {
    EmuScene* es = (EmuScene*)p;
    out[0] = es->hs.n_prims > 1 ? es->hs.n_prims - 1 : 0; out[1] = es->view.n_active_instances > 1 ? es->view.n_active_instances - 1 : 0;
    out[2] = es->blas_depth; out[3] = es->tlas_depth; out[4] = es->view.n_lights; out[5] = es->hs.n_prims; out[6] = es->stack_need; out[7] = es->hs.n_reflective;
    return YRT_OK;
}

// node records (arity 2 or 4) of the emulated build, BLAS nodes then TLAS nodes: returns the float4 count; copies when out != NULL
int64_t emu_read_nodes(void* p, int arity, float* out) {
    EmuScene* es = (EmuScene*)p;
    const std::vector<float4>& v = arity == 4 ? es->nodes4 : es->nodes2;
    // (a tree set of one item has no internal node but the arrays keep one zero record for it: report what the device reports)
    int np = es->hs.n_prims, na = es->view.n_active_instances;
    int64_t n = (int64_t)YRT_NODE_STRIDE(arity) * ((np > 1 ? np - 1 : 0) + (na > 1 ? na - 1 : 0));
    if (out) memcpy(out, v.data(), sizeof(float4) * (size_t)std::min<int64_t>(n, (int64_t)v.size()));
    return n;
}

// counters_out (optional, 8 int64): box tests, prim tests, instance entries, max stack, fused-slab false rejects / extra accepts,
// box tests in the instance tree, node visits
int emu_trace_primary(void* p, const yrt_camera* cam, int width, int height, int samples, int32_t* ids, float* dist, float* uv,
                      int64_t* counters_out) {
    EmuScene* es = (EmuScene*)p;
    camera_k ck = emu_camera(cam);
    EmuGrid cgrid;
    emu_camera_grid(*es, ck, width, height, cgrid);
    const GridRef cg = emu_camera_ref(*es, cgrid);
    long long cb = 0, cp = 0, ci = 0, cfr = 0, cea = 0, ctb = 0, cnv = 0; int cm = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : cb, cp, ci, cfr, cea, ctb, cnv) reduction(max : cm)
    for (int j = 0; j < height; j++) {
        int stack[2 * YRT_STACK_CAP];
        for (int i = 0; i < width; i++)
            for (int jj = 0; jj < samples; jj++)
                for (int ii = 0; ii < samples; ii++) {
                    size_t r = (((size_t)j * width + i) * samples + jj) * samples + ii;
                    float u, v;
                    sample_uv(i, j, ii, jj, samples, width, height, u, v);
                    ray3 ray = eval_camera(ck, u, v);
                    HitRec h;
                    TraceCounters tc = {0, 0, 0, 0, 0, 0, 0, 0};
                    emu_closest(*es, cg, ray, i, j, true, h, stack, &tc);
                    cb += tc.box_tests; cp += tc.prim_tests; ci += tc.inst_entries; cm = std::max(cm, tc.max_stack);
                    cfr += tc.slab_false_rejects; cea += tc.slab_extra_accepts; ctb += tc.tlas_box_tests; cnv += tc.node_visits;
                    hit_to_ids(es->view, h, ids + 3 * r);
                    if (dist) dist[r] = h.dist;
                    if (uv) { uv[2 * r] = h.w1; uv[2 * r + 1] = h.w2; }
                }
    }
    if (counters_out) { counters_out[0] = cb; counters_out[1] = cp; counters_out[2] = ci; counters_out[3] = cm; counters_out[4] = cfr; counters_out[5] = cea; counters_out[6] = ctb; counters_out[7] = cnv; }
    return YRT_OK;
}

// generic rays (8 floats each: o, d, tmin, tmax): closest hit ids/dist + any-hit flags; false_rejects audits the slab test
int emu_intersect(void* p, const float* rays, int64_t n, int32_t* ids, float* dist, uint8_t* occ, int64_t* false_rejects) {
    EmuScene* es = (EmuScene*)p;
    long long fr = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : fr)
    for (int64_t r = 0; r < n; r++) {
        int stack[2 * YRT_STACK_CAP];
        const float* q = rays + 8 * r;
        ray3 ray;
        ray.o = mk3(q[0], q[1], q[2]); ray.d = mk3(q[3], q[4], q[5]); ray.tmin = q[6]; ray.tmax = q[7];
        HitRec h;
        TraceCounters tc = {0, 0, 0, 0, 0, 0, 0, 0};
        emu_closest(*es, gridref_none(), ray, 0, 0, false, h, stack, &tc);
        hit_to_ids(es->view, h, ids + 3 * r);
        dist[r] = h.dist;
        HitRec h2;
        occ[r] = emu_any(*es, ray, h2, stack, &tc) ? 1 : 0;
        fr += tc.slab_false_rejects;
    }
    if (false_rejects) *false_rejects = fr;
    return YRT_OK;
}

// the whole frame with the device functions: raygen -> closest -> shadow (any) -> shade -> reflection
// loop with the same explicit {c, kr, la} stack as k_shade -> ordered per-pixel sum.
// ray_counts (optional, 9 int64): primary, reflection, shadow, then for the shadow rays: box tests, tlas box tests,
// prim tests, instance entries, occluded, node visits
int emu_render(void* p, const yrt_camera* cam, const float amb[3], int width, int height, int samples, int max_depth, float* rgba,
               int64_t* ray_counts) {
    EmuScene* es = (EmuScene*)p;
    const SceneView& sv = es->view;
    camera_k ck = emu_camera(cam);
    EmuGrid cgrid;
    emu_camera_grid(*es, ck, width, height, cgrid);
    const GridRef cg = emu_camera_ref(*es, cgrid);
    vec3 ambv = mk3(amb[0], amb[1], amb[2]);
    if (max_depth <= 0) max_depth = 16;
    long long n_refl = 0, n_shadow = 0, sb = 0, stb = 0, sp_ = 0, si_ = 0, socc = 0, snv = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : n_refl, n_shadow, sb, stb, sp_, si_, socc, snv)
    for (int j = 0; j < height; j++) {
        int stack[2 * YRT_STACK_CAP];
        std::vector<vec3> sc(max_depth), skr(max_depth), sla(max_depth);
        std::vector<uint8_t> vis(std::max(sv.n_lights, 1));
        for (int i = 0; i < width; i++) {
            float sx = 0.f, sy = 0.f, sz = 0.f;
            for (int jj = 0; jj < samples; jj++)
                for (int ii = 0; ii < samples; ii++) {
                    float u, v;
                    sample_uv(i, j, ii, jj, samples, width, height, u, v);
                    ray3 ray = eval_camera(ck, u, v);
                    vec3 value = mk3(0.f, 0.f, 0.f);
                    int depth = 0;
                    for (;;) {
                        HitRec h;
                        emu_closest(*es, cg, ray, i, j, depth == 0, h, stack, nullptr);
                        if (depth > 0) n_refl++;
                        if (h.si < 0) { value = mk3(0.f, 0.f, 0.f); break; }
                        int kind;
                        vec3 P = eval_hit_pos(sv, h.si, h.prim, h.w1, h.w2, kind);
                        for (int k = 0; k < sv.n_lights; k++) {
                            vec3 l, ke; float r;
                            light_vector(sv, k, P, l, r, ke);
                            ray3 sr = shadow_ray(P, l, r);
                            HitRec hr;
                            TraceCounters tc = {0, 0, 0, 0, 0, 0, 0, 0};
                            vis[k] = (es->ref.n_nodes > 0 ? emu_any(*es, sr, hr, stack, &tc) : trace_shadow_ray(sv, es->lg, k, sr, hr, stack, &tc)) ? 0 : 1;
                            n_shadow++;
                            sb += tc.box_tests; stb += tc.tlas_box_tests; sp_ += tc.prim_tests; si_ += tc.inst_entries; socc += vis[k] ? 0 : 1; snv += tc.node_visits;
                        }
                        vec3 c, kr, la; ray3 rr;
                        bool spawn = shade_hit(sv, h.si, h.prim, h.w1, h.w2, ray.o, ambv, sv.srgb_lut, [&](int k) { return vis[k] != 0; },
                                               depth + 1 < max_depth, value, c, kr, la, rr);
                        if (!spawn) break;
                        sc[depth] = c; skr[depth] = kr; sla[depth] = la;
                        ray = rr;
                        depth++;
                    }
                    for (int d = depth - 1; d >= 0; d--) value = combine_reflection(sc[d], value, skr[d], sla[d]);
                    sx += value.x; sy += value.y; sz += value.z;
                }
            float dn = (float)(samples * samples);
            float* o = rgba + 4 * ((size_t)j * width + i);
            o[0] = sx / dn; o[1] = sy / dn; o[2] = sz / dn; o[3] = 1.0f;
        }
    }
    if (ray_counts) { ray_counts[0] = (long long)width * height * samples * samples; ray_counts[1] = n_refl; ray_counts[2] = n_shadow;
        ray_counts[3] = sb; ray_counts[4] = stb; ray_counts[5] = sp_; ray_counts[6] = si_; ray_counts[7] = socc; ray_counts[8] = snv; }
    return YRT_OK;
}

}  // extern "C"
