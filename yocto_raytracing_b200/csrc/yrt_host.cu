// yrt_host.cu — host-only plumbing of libyrt_b200.so: error string, device buffers, and the validated
// host copy of the scene description (everything derived with the HOST libm lives here).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "yrt_internal.h"

namespace yrt {

// ------------------------------------------------------------------------------------------
// errors / buffers
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

int DevBuf::alloc(size_t n, int dev) {
    if (n == 0) n = 16;
    if (p && bytes >= n && device == dev) return YRT_OK;
    release();
    YRT_CUDA(cudaSetDevice(dev));
    YRT_CUDA(cudaMalloc(&p, n));
    bytes = n;
    device = dev;
    return YRT_OK;
}
int DevBuf::upload(const void* src, size_t n, int dev, cudaStream_t st) {
    YRT_TRY(alloc(n, dev));
    if (n) YRT_CUDA(cudaMemcpyAsync(p, src, n, cudaMemcpyHostToDevice, st));
    return YRT_OK;
}
void DevBuf::release() {
    if (p) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (device >= 0 && cur != device) cudaSetDevice(device);
        cudaFree(p);
        if (cur >= 0 && cur != device) cudaSetDevice(cur);
    }
    p = nullptr;
    bytes = 0;
}

// ------------------------------------------------------------------------------------------
// host scene from the C description
// ------------------------------------------------------------------------------------------
int host_scene_from_desc(const yrt_scene_desc* d, HostScene& hs, bool ranks_async) {
    if (!d) { set_error("null scene description"); return YRT_ERR_INVALID; }
    if (d->n_shapes < 0 || d->n_instances < 0 || d->n_materials < 0 || d->n_textures < 0 || d->n_verts < 0 || d->n_elem_idx < 0) {
        set_error("negative count in scene description");
        return YRT_ERR_INVALID;
    }
    if (d->n_shapes >= (1 << 16)) { set_error("more than 65535 shapes (segment id is 16 bits of the sort key)"); return YRT_ERR_UNSUPPORTED; }
    hs.n_shapes = d->n_shapes; hs.n_instances = d->n_instances; hs.n_materials = d->n_materials;
    hs.n_textures = d->n_textures; hs.n_verts = d->n_verts;
#define NEED(ptr, cnt) if ((cnt) > 0 && !(ptr)) { set_error("scene description: %s is null", #ptr); return YRT_ERR_INVALID; }
    NEED(d->shape_kind, d->n_shapes) NEED(d->shape_elem_off, d->n_shapes) NEED(d->shape_elem_cnt, d->n_shapes)
    NEED(d->shape_vert_off, d->n_shapes) NEED(d->shape_vert_cnt, d->n_shapes)
    NEED(d->elem_idx, d->n_elem_idx) NEED(d->pos, d->n_verts) NEED(d->norm, d->n_verts)
    NEED(d->inst_frame, d->n_instances) NEED(d->inst_shape, d->n_instances) NEED(d->inst_mat, d->n_instances)
    NEED(d->mat_ke, d->n_materials) NEED(d->mat_kd, d->n_materials) NEED(d->mat_ks, d->n_materials)
    NEED(d->mat_kr, d->n_materials) NEED(d->mat_rs, d->n_materials) NEED(d->mat_kd_tex, d->n_materials)
    NEED(d->mat_ks_tex, d->n_materials)
    NEED(d->tex_w, d->n_textures) NEED(d->tex_h, d->n_textures) NEED(d->tex_off, d->n_textures) NEED(d->tex_rgba8, d->n_textures)
#undef NEED
    hs.shape_kind.assign(d->shape_kind, d->shape_kind + d->n_shapes);
    hs.shape_elem_off.assign(d->shape_elem_off, d->shape_elem_off + d->n_shapes);
    hs.shape_elem_cnt.assign(d->shape_elem_cnt, d->shape_elem_cnt + d->n_shapes);
    hs.shape_vert_off.assign(d->shape_vert_off, d->shape_vert_off + d->n_shapes);
    hs.shape_vert_cnt.assign(d->shape_vert_cnt, d->shape_vert_cnt + d->n_shapes);
    hs.shape_has_uv.resize(d->n_shapes);
    hs.shape_prim_off.assign(std::max(d->n_shapes, 1) + 1, 0);
    hs.elem_idx.assign(d->elem_idx, d->elem_idx + d->n_elem_idx);
    long long np = 0;
    for (int s = 0; s < d->n_shapes; s++) {
        int kind = hs.shape_kind[s];
        if (kind < 0 || kind > 2) { set_error("shape %d: bad kind %d", s, kind); return YRT_ERR_INVALID; }
        int nv = 3 - kind;
        long long off = hs.shape_elem_off[s], cnt = hs.shape_elem_cnt[s];
        long long vo = hs.shape_vert_off[s], vc = hs.shape_vert_cnt[s];
        if (cnt < 0 || off < 0 || off + cnt * nv > d->n_elem_idx) { set_error("shape %d: element range out of bounds", s); return YRT_ERR_INVALID; }
        if (vc < 0 || vo < 0 || vo + vc > d->n_verts) { set_error("shape %d: vertex range out of bounds", s); return YRT_ERR_INVALID; }
        for (long long k = off; k < off + cnt * nv; k++)
            if (hs.elem_idx[k] < 0 || hs.elem_idx[k] >= vc) { set_error("shape %d: vertex index %d out of range [0,%lld)", s, hs.elem_idx[k], vc); return YRT_ERR_INVALID; }
        hs.shape_has_uv[s] = (d->shape_has_uv && d->uv) ? (d->shape_has_uv[s] != 0) : (d->uv != nullptr && !d->shape_has_uv);
        if (kind != YRT_TRIANGLES && cnt > 0 && (!d->radius || (d->shape_has_radius && !d->shape_has_radius[s]))) {
            set_error("shape %d: lines/points need per-vertex radius (the reference indexes shape::radius unconditionally, scene.cpp:419,431)", s);
            return YRT_ERR_INVALID;
        }
        hs.shape_prim_off[s] = (int)np;
        np += cnt;
    }
    if (np > YRT_MAX_LEAF_FIRST) { set_error("too many elements (%lld)", np); return YRT_ERR_UNSUPPORTED; }
    hs.shape_prim_off[d->n_shapes] = (int)np;
    hs.n_prims = (int)np;
    hs.prim_shape.resize(hs.n_prims);
    for (int s = 0; s < d->n_shapes; s++)
        for (int k = hs.shape_prim_off[s]; k < hs.shape_prim_off[s + 1]; k++) hs.prim_shape[k] = s;
    hs.pos.assign(d->pos, d->pos + 3 * (size_t)d->n_verts);
    hs.norm.assign(d->norm, d->norm + 3 * (size_t)d->n_verts);
    if (d->uv) hs.uv.assign(d->uv, d->uv + 2 * (size_t)d->n_verts); else hs.uv.assign(2 * (size_t)d->n_verts, 0.f);
    if (d->radius) hs.radius.assign(d->radius, d->radius + (size_t)d->n_verts); else hs.radius.assign((size_t)d->n_verts, 0.f);

    hs.inst_frame.assign(d->inst_frame, d->inst_frame + 12 * (size_t)d->n_instances);
    hs.inst_shape.assign(d->inst_shape, d->inst_shape + d->n_instances);
    hs.inst_mat.assign(d->inst_mat, d->inst_mat + d->n_instances);
    hs.active_inst.clear();
    for (int i = 0; i < d->n_instances; i++) {
        if (hs.inst_shape[i] < 0 || hs.inst_shape[i] >= d->n_shapes) { set_error("instance %d: bad shape index", i); return YRT_ERR_INVALID; }
        if (hs.inst_mat[i] < 0 || hs.inst_mat[i] >= d->n_materials) { set_error("instance %d: bad material index", i); return YRT_ERR_INVALID; }
        if (hs.shape_elem_cnt[hs.inst_shape[i]] > 0) hs.active_inst.push_back(i);
    }
    if ((long long)hs.active_inst.size() > YRT_MAX_LEAF_FIRST) { set_error("too many instances"); return YRT_ERR_UNSUPPORTED; }

    // textures
    hs.tex_info.resize(d->n_textures);
    for (int t = 0; t < d->n_textures; t++) {
        long long w = d->tex_w[t], h = d->tex_h[t], off = d->tex_off[t];
        if (w <= 0 || h <= 0) { set_error("texture %d is empty (the reference would index a 0x0 image, raytrace.cpp:43)", t); return YRT_ERR_INVALID; }
        if (off < 0 || off + 4 * w * h > d->tex_bytes) { set_error("texture %d: texel range out of bounds", t); return YRT_ERR_INVALID; }
        int4 ti; ti.x = (int)w; ti.y = (int)h; ti.z = (int)(unsigned)(off & 0xffffffffll); ti.w = (int)(unsigned)((unsigned long long)off >> 32);
        hs.tex_info[t] = ti;
    }
    hs.tex_rgba8.assign(d->tex_rgba8, d->tex_rgba8 + (d->n_textures ? d->tex_bytes : 0));
    // lookup_texture's gamma decode (raytrace.cpp:51-53) through the HOST libm, once per byte value
    for (int b = 0; b < 256; b++) hs.srgb_lut[b] = fminf(1.0f, powf((float)b / 255.0f, 2.2f));

    // materials; ns as in raytrace.cpp:144 with the host powf
    hs.mat_recs.resize(4 * (size_t)d->n_materials);
    hs.n_reflective = 0;
    for (int m = 0; m < d->n_materials; m++) {
        float rs = d->mat_rs[m];
        float ns = (rs) ? 2 / powf(rs, 4.0f) - 2 : 1e6f;
        int kdt = d->mat_kd_tex[m], kst = d->mat_ks_tex[m];
        if (kdt >= d->n_textures || kst >= d->n_textures) { set_error("material %d: bad texture index", m); return YRT_ERR_INVALID; }
        if (kdt < 0) kdt = -1;
        if (kst < 0) kst = -1;
        const float *kd = d->mat_kd + 3 * m, *ks = d->mat_ks + 3 * m, *kr = d->mat_kr + 3 * m, *ke = d->mat_ke + 3 * m;
        hs.mat_recs[4 * m + 0] = mk4(kd[0], kd[1], kd[2], ns);
        hs.mat_recs[4 * m + 1] = mk4(ks[0], ks[1], ks[2], int_as_float(kdt));
        hs.mat_recs[4 * m + 2] = mk4(kr[0], kr[1], kr[2], int_as_float(kst));
        hs.mat_recs[4 * m + 3] = mk4(ke[0], ke[1], ke[2], 0.f);
        if (kr[0] > 0.0f || kr[1] > 0.0f || kr[2] > 0.0f) hs.n_reflective++;
    }
    // lights: every instance whose material has ke.x>0 && ke.y>0 && ke.z>0 (raytrace.cpp:126), in
    // instance order; position = shape::pos.front() (raytrace.cpp:129)
    hs.light_recs.clear();
    hs.light_inst.clear();
    for (int i = 0; i < d->n_instances; i++) {
        const float* ke = d->mat_ke + 3 * hs.inst_mat[i];
        if (!(ke[0] > 0.0f && ke[1] > 0.0f && ke[2] > 0.0f)) continue;
        int s = hs.inst_shape[i];
        if (hs.shape_vert_cnt[s] <= 0) { set_error("light instance %d has a shape without vertices (pos.front() is undefined)", i); return YRT_ERR_INVALID; }
        const float* fr = d->inst_frame + 12 * (size_t)i;
        const float* p0 = d->pos + 3 * (size_t)hs.shape_vert_off[s];
        hs.light_recs.push_back(mk4(fr[0], fr[1], fr[2], ke[0]));
        hs.light_recs.push_back(mk4(fr[3], fr[4], fr[5], ke[1]));
        hs.light_recs.push_back(mk4(fr[6], fr[7], fr[8], ke[2]));
        hs.light_recs.push_back(mk4(fr[9], fr[10], fr[11], 0.f));
        hs.light_recs.push_back(mk4(p0[0], p0[1], p0[2], 0.f));
        hs.light_inst.push_back(i);
    }
    // bound of every world coordinate (error budget of the apex grids) and rigidity of the instance frames
    {
        std::vector<float> shape_r(std::max(hs.n_shapes, 1), 0.f);
        for (int s = 0; s < hs.n_shapes; s++) {
            float m = 0.f;
            for (int v = hs.shape_vert_off[s]; v < hs.shape_vert_off[s] + hs.shape_vert_cnt[s]; v++) {
                float rad = hs.radius.empty() ? 0.f : fabsf(hs.radius[v]);
                for (int c = 0; c < 3; c++) m = std::max(m, fabsf(hs.pos[3 * (size_t)v + c]) + rad);
            }
            shape_r[s] = m;
        }
        hs.extent = 0.f; hs.all_rigid = true;
        for (int i = 0; i < hs.n_instances; i++) {
            const float* f = &hs.inst_frame[12 * (size_t)i];
            float rows = 0.f;
            for (int c = 0; c < 3; c++) rows = std::max(rows, fabsf(f[c]) + fabsf(f[3 + c]) + fabsf(f[6 + c]));
            float e = std::max(std::max(fabsf(f[9]), fabsf(f[10])), fabsf(f[11])) + rows * shape_r[hs.inst_shape[i]];
            if (e > hs.extent || !(e == e)) hs.extent = e;
            auto dot3 = [](const float* a, const float* b) { return (double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2]; };
            const double tol = 1e-4;
            bool rigid = fabs(dot3(f, f) - 1.0) <= tol && fabs(dot3(f + 3, f + 3) - 1.0) <= tol && fabs(dot3(f + 6, f + 6) - 1.0) <= tol &&
                         fabs(dot3(f, f + 3)) <= tol && fabs(dot3(f, f + 6)) <= tol && fabs(dot3(f + 3, f + 6)) <= tol;
            if (!rigid) hs.all_rigid = false;
        }
    }
    if (ranks_async) hs.rank_thread = std::thread([&hs]() { reference_visit_ranks(hs); });
    else reference_visit_ranks(hs);
    return YRT_OK;
}

// ------------------------------------------------------------------------------------------
// Tie-breaking order.  intersect_triangle/line/point reject only t > tmax (src/scene.cpp:256,271,295),
// so a candidate at EXACTLY the current closest distance replaces it: among exact ties the reference
// returns the one its traversal visits last.  Its traversal order does not depend on the ray: nodes
// push both children and pop the second first (scene.cpp:402-404,462-464), leaves run in array order.
// So the tie winner is fixed by the reference's tree alone.  instance10000 has many such ties
// (neighbouring instances overlap and share the plane y = 2), and the two candidates carry different
// materials, so image parity needs the same winner.  This restates ONLY the partition structure of
// build_bvh/make_node/split_prims (scene.cpp:525-658) — no nodes, no boxes are kept — and records, per
// element and per instance, its position in that visit sequence.  The traversal BVH itself is the GPU
// LBVH; these ranks are consulted on exact ties only (yrt_trace.cuh).
// ------------------------------------------------------------------------------------------
namespace {
struct RankItem { vec3 center; int pid; Box box; };

// std::partition as libstdc++ implements it for bidirectional iterators (the reference is compiled
// against libstdc++ here; the element order inside each side depends on this exact algorithm)
template <class Pred> int partition_like_libstdcxx(std::vector<RankItem>& v, int first, int last, Pred pred) {
    for (;;) {
        for (;;) {
            if (first == last) return first;
            if (pred(v[first])) ++first; else break;
        }
        --last;
        for (;;) {
            if (first == last) return first;
            if (!pred(v[last])) --last; else break;
        }
        std::swap(v[first], v[last]);
        ++first;
    }
}

// ranks[pid] = position of pid in the visit sequence of the tree make_node() builds over items.
// ref_nodes / ref_leaf (optional): the tree itself, for scenes whose result depends on it (RefTlas, yrt_scene.cuh) —
// per node two quads {bbox.min | a}, {bbox.max | b}: inner node a = left child, b = right child (the reference pushes
// `start`, `start + 1` and pops the second first, scene.cpp:462-464); leaf a = ~(first entry of ref_leaf), b = count;
// ref_leaf = the pids in partition order (bvh->leaf_prims, scene.cpp:652-655).  A node's box is the union of its items'
// boxes in array order (scene.cpp:573-576).
void visit_ranks(std::vector<RankItem>& items, std::vector<int>& ranks, int rank_base_index, std::vector<float4>* ref_nodes = nullptr,
                 std::vector<int>* ref_leaf = nullptr) {
    struct Range { int start, end; bool expanded; };
    // phase 1: partition exactly like make_node (pre-order: node, then left subtree, then right subtree;
    // the two recursions are independent, so only the split points matter); remember the tree
    struct Node { int start, end, left, right; };
    std::vector<Node> nodes;
    nodes.push_back({0, (int)items.size(), -1, -1});
    std::vector<int> todo(1, 0);
    while (!todo.empty()) {
        int nid = todo.back();
        todo.pop_back();
        int start = nodes[nid].start, end = nodes[nid].end;
        if (end - start <= 4) continue;                                   // scene.cpp:583
        vec3 lo = mk3(FLT_MAX, FLT_MAX, FLT_MAX), hi = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        for (int i = start; i < end; i++) {                               // scene.cpp:609-611
            const vec3& c = items[i].center;
            lo = mk3(rmin(lo.x, c.x), rmin(lo.y, c.y), rmin(lo.z, c.z));
            hi = mk3(rmax(hi.x, c.x), rmax(hi.y, c.y), rmax(hi.z, c.z));
        }
        vec3 size = hi - lo;
        if (size.x == 0 && size.y == 0 && size.z == 0) continue;          // scene.cpp:614
        int axis;
        if (size.x >= size.y && size.x >= size.z) axis = 0;               // scene.cpp:616-621
        else if (size.y >= size.x && size.y >= size.z) axis = 1;
        else axis = 2;
        vec3 half = (lo + hi) / 2;                                        // scene.cpp:630
        float h = axis == 0 ? half.x : (axis == 1 ? half.y : half.z);
        int mid = partition_like_libstdcxx(items, start, end, [axis, h](const RankItem& a) {
            float c = axis == 0 ? a.center.x : (axis == 1 ? a.center.y : a.center.z);
            return c < h;
        });
        if (mid <= start || mid >= end) continue;   // the reference asserts (compiled out) and would recurse forever
        int l = (int)nodes.size();
        nodes.push_back({start, mid, -1, -1});
        nodes.push_back({mid, end, -1, -1});
        nodes[nid].left = l;
        nodes[nid].right = l + 1;
        todo.push_back(l);
        todo.push_back(l + 1);
    }
    if (ref_nodes && ref_leaf) {
        ref_nodes->assign(2 * nodes.size(), mk4(0.f, 0.f, 0.f, 0.f));
        ref_leaf->resize(items.size());
        for (size_t i = 0; i < items.size(); i++) (*ref_leaf)[i] = items[i].pid;
        for (size_t k = 0; k < nodes.size(); k++) {
            const Node& n = nodes[k];
            Box b = box_invalid();
            for (int i = n.start; i < n.end; i++) box_expand(b, items[i].box.lo, items[i].box.hi);   // expand_bbox(node->bbox, leaf_prims[i].bbox)
            const int qa = n.left >= 0 ? n.left : ~n.start, qb = n.left >= 0 ? n.right : n.end - n.start;
            (*ref_nodes)[2 * k] = mk4(b.lo.x, b.lo.y, b.lo.z, int_as_float(qa));
            (*ref_nodes)[2 * k + 1] = mk4(b.hi.x, b.hi.y, b.hi.z, int_as_float(qb));
        }
    }
    // phase 2: visit order = stack traversal that pushes (first, first+1) and pops first+1 first
    int counter = 0;
    std::vector<int> st(1, 0);
    while (!st.empty()) {
        int nid = st.back();
        st.pop_back();
        const Node& n = nodes[nid];
        if (n.left < 0) {
            for (int i = n.start; i < n.end; i++) ranks[rank_base_index + items[i].pid] = counter++;
        } else {
            st.push_back(n.left);
            st.push_back(n.right);
        }
    }
}
}  // namespace

void reference_visit_ranks(HostScene& hs) {
    GeomView g;
    g.shape_kind = hs.shape_kind.data(); g.shape_elem_off = hs.shape_elem_off.data(); g.shape_elem_cnt = hs.shape_elem_cnt.data();
    g.shape_vert_off = hs.shape_vert_off.data(); g.shape_prim_off = hs.shape_prim_off.data(); g.elem_idx = hs.elem_idx.data();
    g.pos = hs.pos.data(); g.norm = hs.norm.data(); g.uv = hs.uv.data(); g.radius = hs.radius.data();
    g.prim_shape = hs.prim_shape.data(); g.n_prims = hs.n_prims;
    hs.prim_rank.assign(std::max(hs.n_prims, 1), 0);
    hs.inst_rank.assign(std::max(hs.n_instances, 1), 0);
    std::vector<Box> shape_box(std::max(hs.n_shapes, 1), box_invalid());

    std::vector<RankItem> items;
    for (int s = 0; s < hs.n_shapes; s++) {
        int p0 = hs.shape_prim_off[s], p1 = hs.shape_prim_off[s + 1];
        items.clear();
        Box sb = box_invalid();
        for (int gp = p0; gp < p1; gp++) {
            Box b = prim_bounds(g, gp);                                   // scene.cpp:527-547
            RankItem it;
            it.center = (b.lo + b.hi) / 2.0f;
            it.pid = gp - p0;
            it.box = b;
            items.push_back(it);
            box_expand(sb, b.lo, b.hi);
        }
        shape_box[s] = sb;                                                // bvh->nodes[0].bbox
        visit_ranks(items, hs.prim_rank, p0);
    }
    items.clear();
    for (int i = 0; i < hs.n_instances; i++) {                            // scene.cpp:558-562
        const float* fr = &hs.inst_frame[12 * (size_t)i];
        frame3 f;
        f.x = mk3(fr[0], fr[1], fr[2]); f.y = mk3(fr[3], fr[4], fr[5]); f.z = mk3(fr[6], fr[7], fr[8]); f.o = mk3(fr[9], fr[10], fr[11]);
        Box w = instance_bounds(f, shape_box[hs.inst_shape[i]]);
        RankItem it;
        it.center = (w.lo + w.hi) / 2.0f;
        it.pid = i;
        it.box = w;
        items.push_back(it);
    }
    hs.ref_nodes.clear();
    hs.ref_leaf_inst.clear();
    // scaled / sheared instance frames: the rays walk the reference's own instance tree, in its order (see RefTlas, yrt_scene.cuh)
    if (hs.all_rigid) visit_ranks(items, hs.inst_rank, 0);
    else visit_ranks(items, hs.inst_rank, 0, &hs.ref_nodes, &hs.ref_leaf_inst);
}

}  // namespace yrt
