// yrt_build.cu — scene upload and the GPU two-level LBVH build (sm_100a).
//
// Replaces build_bvh(scn,false) (src/raytrace.cpp:278 -> src/scene.cpp:554-565 and :525-549,
// :572-658).  Launch sequence per tree set (BLAS over all shapes' elements at once, then TLAS
// over instances): bounds -> per-segment reductions -> Morton keys -> 8 x {histogram, scan,
// stable scatter} LSD radix sort -> Karras topology -> atomic bottom-up refit -> node emit.
#include <algorithm>
#include <chrono>
#include <mutex>
#include <utility>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "yrt_internal.h"

namespace yrt {

// ------------------------------------------------------------------------------------------
// kernels: one thread per item, bodies in yrt_lbvh.cuh
// ------------------------------------------------------------------------------------------
#define YRT_TID() ((int)(blockIdx.x * blockDim.x + threadIdx.x))
static inline int grid_for(int n, int t = 256) { return n > 0 ? (n + t - 1) / t : 1; }
static inline int grid_for_warps(int n_warps) { return n_warps > 0 ? (n_warps + 7) / 8 : 1; }   // 256-thread blocks, one warp per item

__global__ void k_prim_boxes(GeomView g, float4* lo, float4* hi) {
    int i = YRT_TID();
    if (i >= g.n_prims) return;
    Box b = prim_bounds(g, i);
    lo[i] = mk4(b.lo.x, b.lo.y, b.lo.z, 0.f);
    hi[i] = mk4(b.hi.x, b.hi.y, b.hi.z, 0.f);
}
__global__ void k_seg_init(LbvhArrays a) { int s = YRT_TID(); if (s < a.n_seg) seg_bounds_init_item(a, s); }
__global__ void k_seg_bounds(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) seg_bounds_item(a, i); }
__global__ void k_morton(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) morton_item(a, i); }
__global__ void k_karras(LbvhArrays a) { int i = YRT_TID(); if (i < a.n - 1) karras_item(a, i); }
__global__ void k_refit(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) refit_item(a, i); }
__global__ void k_emit(LbvhArrays a) { int i = YRT_TID(); if (i < a.n - 1) emit_item(a, i); }
__global__ void k_rotate_refit(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) rotate_refit_item(a, i); }
__global__ void k_relayout_slot(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) relayout_slot_item(a, i); }
__global__ void k_relayout_move(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) relayout_move_item(a, i); }
__global__ void k_relayout_refs(LbvhArrays a) { int i = YRT_TID(); if (i < a.n - 1) relayout_refs_item(a, i); }
__global__ void k_relayout_copy(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) relayout_copy_item(a, i); }
__global__ void k_single_root(LbvhArrays a) { int s = YRT_TID(); if (s < a.n_seg) single_root_item(a, s); }
__global__ void k_depth(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) depth_item(a, i); }
__global__ void k_stackneed(LbvhArrays a) { int i = YRT_TID(); if (i < a.n) stackneed_item(a, i); }

// prim + attribute records in BLAS leaf order
__global__ void k_gather_prims(GeomView g, const int* __restrict__ order, float4* __restrict__ prim_recs,
                               float4* __restrict__ prim_attrs, const int* __restrict__ shape_has_uv) {
    int k = YRT_TID();
    if (k >= g.n_prims) return;
    int gp = order[k];
    int s = g.prim_shape[gp];
    int e = gp - g.shape_prim_off[s];
    int kind = g.shape_kind[s];
    int vo = g.shape_vert_off[s];
    int nv = kind == 0 ? 3 : (kind == 1 ? 2 : 1);
    const int* t = g.elem_idx + g.shape_elem_off[s] + (size_t)nv * e;
    int v[3] = {vo + t[0], vo + t[nv > 1 ? 1 : 0], vo + t[nv > 2 ? 2 : 0]};
    vec3 p0 = ld3(g.pos, v[0]), p1 = ld3(g.pos, v[1]), p2 = ld3(g.pos, v[2]);
    vec3 n0 = ld3(g.norm, v[0]), n1 = ld3(g.norm, v[1]), n2 = ld3(g.norm, v[2]);
    bool huv = shape_has_uv[s] != 0;
    float u0 = huv ? g.uv[2 * (size_t)v[0]] : 0.f, w0 = huv ? g.uv[2 * (size_t)v[0] + 1] : 0.f;
    float u1 = huv ? g.uv[2 * (size_t)v[1]] : 0.f, w1 = huv ? g.uv[2 * (size_t)v[1] + 1] : 0.f;
    float u2 = huv ? g.uv[2 * (size_t)v[2]] : 0.f, w2 = huv ? g.uv[2 * (size_t)v[2] + 1] : 0.f;
    float4* pr = prim_recs + 3 * (size_t)k;
    float4* ar = prim_attrs + YRT_ATTR_STRIDE * (size_t)k;
    pack_prim(kind, e, p0, p1, p2, kind == 0 ? 0.f : g.radius[v[0]], kind == 1 ? g.radius[v[1]] : 0.f, pr, ar);
    ar[0] = mk4(n0.x, n0.y, n0.z, u0);
    ar[1] = mk4(n1.x, n1.y, n1.z, w0);
    ar[2] = mk4(n2.x, n2.y, n2.z, u1);
    ar[3] = mk4(w1, u2, w2, 0.f);
}

// world bounds of every active instance (TLAS leaves)
__global__ void k_inst_boxes(int n_active, const int* __restrict__ active_inst, const float* __restrict__ inst_frame,
                             const int* __restrict__ inst_shape, const int* __restrict__ seg_box_lo,
                             const int* __restrict__ seg_box_hi, float4* lo, float4* hi) {
    int a = YRT_TID();
    if (a >= n_active) return;
    int inst = active_inst[a];
    int s = inst_shape[inst];
    const float* fr = inst_frame + 12 * (size_t)inst;
    frame3 f;
    f.x = mk3(fr[0], fr[1], fr[2]); f.y = mk3(fr[3], fr[4], fr[5]); f.z = mk3(fr[6], fr[7], fr[8]); f.o = mk3(fr[9], fr[10], fr[11]);
    Box b;
    b.lo = mk3(ordered_to_float(seg_box_lo[3 * s]), ordered_to_float(seg_box_lo[3 * s + 1]), ordered_to_float(seg_box_lo[3 * s + 2]));
    b.hi = mk3(ordered_to_float(seg_box_hi[3 * s]), ordered_to_float(seg_box_hi[3 * s + 1]), ordered_to_float(seg_box_hi[3 * s + 2]));
    Box w = instance_bounds(f, b);
    lo[a] = mk4(w.lo.x, w.lo.y, w.lo.z, 0.f);
    hi[a] = mk4(w.hi.x, w.hi.y, w.hi.z, 0.f);
}

// instance records in TLAS leaf order
__global__ void k_inst_recs(int n_active, const int* __restrict__ order, const int* __restrict__ active_inst,
                            const float* __restrict__ inst_frame, const int* __restrict__ inst_shape,
                            const int* __restrict__ inst_mat, const int* __restrict__ shape_kind,
                            const int* __restrict__ blas_root, float4* __restrict__ recs) {
    int k = YRT_TID();
    if (k >= n_active) return;
    int inst = active_inst[order[k]];
    int s = inst_shape[inst];
    const float* fr = inst_frame + 12 * (size_t)inst;
    float4* r = recs + 4 * (size_t)k;
    r[0] = mk4(fr[0], fr[1], fr[2], int_as_float(blas_root[s]));
    r[1] = mk4(fr[3], fr[4], fr[5], int_as_float(inst));
    r[2] = mk4(fr[6], fr[7], fr[8], int_as_float(inst_mat[inst]));
    r[3] = mk4(fr[9], fr[10], fr[11], int_as_float((int)((unsigned)s | ((unsigned)shape_kind[s] << 28))));
}

// world boxes of the instances in TLAS leaf order, as (centre, inflated half-extent): the boxes the apex grids are built from
__global__ void k_inst_box(int n_active, const int* __restrict__ order, const float4* __restrict__ lo, const float4* __restrict__ hi,
                           float4* __restrict__ out) {
    int k = YRT_TID();
    if (k >= n_active) return;
    const float4 l = lo[order[k]], h = hi[order[k]];
    float cx, cy, cz, hx, hy, hz;
    box_center_half(l.x, h.x, cx, hx);
    box_center_half(l.y, h.y, cy, hy);
    box_center_half(l.z, h.z, cz, hz);
    out[2 * (size_t)k] = mk4(cx, cy, cz, 0.f);
    out[2 * (size_t)k + 1] = mk4(hx, hy, hz, 0.f);
}

// ------------------------------------------------------------------------------------------
// apex grids (yrt_pgrid.cuh): count -> allocate -> fill -> sort, per grid
// ------------------------------------------------------------------------------------------
#define PGRID_BIG_BLOCKS 64
__global__ void __launch_bounds__(256) k_pgrid_scatter(PGridArrays a, int fill) {   // one warp per instance slot
    const int w = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w < a.n_inst) pgrid_scatter_item(a, w, lane, 32, fill != 0, false);
}
__global__ void __launch_bounds__(256) k_pgrid_scatter_big(PGridArrays a, int fill) {   // all threads of the grid per large instance
    const int n_big = a.big[0], t = YRT_TID(), nt = (int)(gridDim.x * blockDim.x);
    for (int b = 0; b < n_big; b++) pgrid_scatter_item(a, a.big[1 + b], t, nt, fill != 0, true);
}
__global__ void __launch_bounds__(256) k_pgrid_alloc(PGridArrays a) {
    const int cell = YRT_TID(), lane = threadIdx.x & 31;
    const bool ok = cell < a.d.n_cells;
    const int n = ok ? a.cnt[cell] : 0;
    if (ok) a.cnt[cell] = 0;
    const int need = pgrid_alloc_need(a, n);
    int x = need;   // inclusive prefix over the warp: one atomic on the running total per warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    const int tot = __shfl_sync(0xffffffffu, x, 31);
    int base = 0;
    if (lane == 0 && tot) base = atomicAdd(a.total, tot);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) a.cells[cell] = pgrid_alloc_cell(a, n, base + x - need);
}
__global__ void __launch_bounds__(128) k_pgrid_emit(PGridArrays a) { int c = YRT_TID(); if (c < a.d.n_cells) pgrid_emit_item(a, c); }
__global__ void __launch_bounds__(256) k_pgrid_emit_slots(PGridArrays a) { int t = YRT_TID(); if (t < a.d.n_cells * a.fixed_k) pgrid_emit_slot_item(a, t); }

int pgrid_build_enqueue(const PGridArrays& a, cudaStream_t st) {
    if (a.d.mode == 0 || a.d.n_cells <= 0) return YRT_OK;
    // cnt | total | big[0] are one allocation (see the callers): cleared together (a.node_total is the caller's: grids may share a pool)
    YRT_CUDA(cudaMemsetAsync(a.cnt, 0, sizeof(int) * ((size_t)a.d.n_cells + 2), st));
    const int warps_grid = grid_for_warps(a.n_inst);
    if (a.as_list && a.fixed_k > 0) {
        // one scatter pass into the cells' own slots, then one thread per (cell, slot)
        if (a.n_inst > 0) {
            k_pgrid_scatter<<<warps_grid, 256, 0, st>>>(a, 0);
            k_pgrid_scatter_big<<<PGRID_BIG_BLOCKS, 256, 0, st>>>(a, 0);
        }
        k_pgrid_emit_slots<<<(a.d.n_cells * a.fixed_k + 255) / 256, 256, 0, st>>>(a);
        YRT_CUDA(cudaGetLastError());
        return YRT_OK;
    }
    if (a.n_inst > 0) {
        k_pgrid_scatter<<<warps_grid, 256, 0, st>>>(a, 0);
        k_pgrid_scatter_big<<<PGRID_BIG_BLOCKS, 256, 0, st>>>(a, 0);
    }
    k_pgrid_alloc<<<(a.d.n_cells + 255) / 256, 256, 0, st>>>(a);
    if (a.n_inst > 0) {
        k_pgrid_scatter<<<warps_grid, 256, 0, st>>>(a, 1);
        k_pgrid_scatter_big<<<PGRID_BIG_BLOCKS, 256, 0, st>>>(a, 1);
    }
    k_pgrid_emit<<<(a.d.n_cells + 127) / 128, 128, 0, st>>>(a);
    YRT_CUDA(cudaGetLastError());
    return YRT_OK;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, 64-bit keys + int values, multi-CTA
// ------------------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 4
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned long long* __restrict__ keys, int n, int shift,
                                                        int* __restrict__ ghist, int nblocks) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int i = base + it * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 0xff)], 1);
    }
    __syncthreads();
    ghist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of ghist[0 .. len) in place, single block
__global__ void __launch_bounds__(1024) k_rs_scan(int* __restrict__ ghist, int len) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < len; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < len ? ghist[i] : 0;
        int x = v;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[w] = x;
        __syncthreads();
        if (w == 0) {
            int s = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        int carry = carry_s;
        int incl = x + (w > 0 ? warp_sums[w - 1] : 0);
        if (i < len) ghist[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned long long* __restrict__ keys_in,
                                                           const int* __restrict__ vals_in,
                                                           unsigned long long* __restrict__ keys_out,
                                                           int* __restrict__ vals_out, int n, int shift,
                                                           const int* __restrict__ gscan, int nblocks) {
    __shared__ int wh[RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int base = blockIdx.x * RS_TILE + w * (32 * RS_ITEMS);
    unsigned long long key[RS_ITEMS];
    int val[RS_ITEMS], rank[RS_ITEMS], dig[RS_ITEMS];
    unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int i = base + it * 32 + lane;
        bool ok = i < n;
        key[it] = ok ? keys_in[i] : 0ull;
        val[it] = ok ? vals_in[i] : 0;
        int d = ok ? (int)((key[it] >> shift) & 0xff) : 256;   // 256 = padding lanes, never ranked
        dig[it] = d;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int prior = 0;
        if (ok) prior = wh[w][d];
        __syncwarp();
        if (ok && (peers & lt) == 0) wh[w][d] = prior + __popc(peers);
        __syncwarp();
        rank[it] = prior + __popc(peers & lt);
    }
    __syncthreads();
    {   // per digit: exclusive prefix over the warps of this block
        int d = threadIdx.x;
        int run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            int t = wh[ww][d];
            wh[ww][d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        int i = base + it * 32 + lane;
        if (i < n) {
            int d = dig[it];
            int dst = gscan[d * nblocks + blockIdx.x] + wh[w][d] + rank[it];
            keys_out[dst] = key[it];
            vals_out[dst] = val[it];
        }
    }
}

static int radix_sort_pairs(cudaStream_t st, unsigned long long*& keys, int*& vals, unsigned long long*& keys_alt,
                            int*& vals_alt, int n, int* ghist) {
    if (n <= 1) return YRT_OK;
    int nblocks = (n + RS_TILE - 1) / RS_TILE;
    for (int pass = 0; pass < 8; pass++) {
        int shift = pass * 8;
        k_rs_hist<<<nblocks, RS_THREADS, 0, st>>>(keys, n, shift, ghist, nblocks);
        k_rs_scan<<<1, 1024, 0, st>>>(ghist, 256 * nblocks);
        k_rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(keys, vals, keys_alt, vals_alt, n, shift, ghist, nblocks);
        std::swap(keys, keys_alt);
        std::swap(vals, vals_alt);
    }
    YRT_CUDA(cudaGetLastError());
    return YRT_OK;
}
static size_t radix_hist_ints(int n) { return 256 * (size_t)((std::max(n, 1) + RS_TILE - 1) / RS_TILE); }

// test hook declared in include/yrt_b200.h: sort n (key, value) pairs held in HOST arrays on the device
extern "C" int yrt_debug_sort_pairs(unsigned long long* h_keys, int* h_vals, int n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return YRT_ERR_NO_DEVICE; }
    DevBuf k0, k1, v0, v1, gh;
    YRT_TRY(k0.upload(h_keys, sizeof(unsigned long long) * (size_t)n, dev, 0));
    YRT_TRY(v0.upload(h_vals, sizeof(int) * (size_t)n, dev, 0));
    YRT_TRY(k1.alloc(sizeof(unsigned long long) * (size_t)n, dev));
    YRT_TRY(v1.alloc(sizeof(int) * (size_t)n, dev));
    YRT_TRY(gh.alloc(sizeof(int) * radix_hist_ints(n), dev));
    unsigned long long *ka = k0.as<unsigned long long>(), *kb = k1.as<unsigned long long>();
    int *va = v0.as<int>(), *vb = v1.as<int>();
    YRT_TRY(radix_sort_pairs(0, ka, va, kb, vb, n, gh.as<int>()));
    YRT_CUDA(cudaMemcpy(h_keys, ka, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToHost));
    YRT_CUDA(cudaMemcpy(h_vals, va, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    return YRT_OK;
}

// ------------------------------------------------------------------------------------------
// device memory of one scene: two arenas (what stays / what the build needs), each ONE cudaMalloc.
// An Arena is walked twice with the same sequence of take() calls: a dry pass that only adds up sizes, then the real
// one — so the layout is written down once, in the order the build uses it.
// ------------------------------------------------------------------------------------------
struct Arena {
    char* base = nullptr;
    size_t off = 0;
    template <class T> T* take(size_t count) {
        size_t a = (off + 255) & ~(size_t)255;
        off = a + std::max(sizeof(T) * count, (size_t)16);
        return base ? (T*)(base + a) : nullptr;
    }
};

// ------------------------------------------------------------------------------------------
// one LBVH build (many segments): enqueues everything on `st`, never synchronises
// ------------------------------------------------------------------------------------------
struct LbvhIo {
    // in
    int n, n_seg, leaf_size, ref_offset, size_bits, rotate_rounds, rotate_pairs;
    float4 *box_lo, *box_hi;       // [n] item bounds
    const int *seg_of, *seg_first;
    float4 *nodes2, *nodes4;       // this tree set's records in the two shared node arrays (binary: 4, 4-wide: 8 float4 per node)
    // out (persistent arena)
    int *seg_root, *seg_depth, *seg_need2, *seg_need4;   // [n_seg]
    int *seg_box_lo, *seg_box_hi;           // [3*n_seg] ordered ints
    int* order;                             // [n] item id at sorted slot (temp arena, read by the gather kernels)
};



// temporaries of one build from `tmp` (dry or real), then — if real — the launches
static int lbvh_build(cudaStream_t st, Arena& tmp, LbvhIo& io) {
    const int n = io.n, n_seg = io.n_seg;
    const size_t nn = (size_t)std::max(n, 1), ni = n > 1 ? (size_t)(n - 1) : 1;
    LbvhArrays a;
    a.n = n; a.n_seg = n_seg;
    a.box_lo = io.box_lo; a.box_hi = io.box_hi; a.seg_of = io.seg_of; a.seg_first = io.seg_first;
    a.seg_cent_lo = tmp.take<int>(3 * (size_t)n_seg);
    a.seg_cent_hi = tmp.take<int>(3 * (size_t)n_seg);
    a.seg_box_lo = io.seg_box_lo; a.seg_box_hi = io.seg_box_hi;
    a.keys = tmp.take<unsigned long long>(nn);
    unsigned long long* keys_alt = tmp.take<unsigned long long>(nn);
    io.order = a.order = tmp.take<int>(nn);
    int* order_alt = tmp.take<int>(nn);
    a.left = tmp.take<int>(ni); a.right = tmp.take<int>(ni);
    a.range_first = tmp.take<int>(ni); a.range_last = tmp.take<int>(ni);
    a.parent_int = tmp.take<int>(ni); a.parent_leaf = tmp.take<int>(nn);
    a.flags = tmp.take<int>(ni);
    a.node_lo = tmp.take<float4>(ni); a.node_hi = tmp.take<float4>(ni);
    a.count = tmp.take<int>(ni); a.new_slot = tmp.take<int>(nn);
    a.order_tmp = tmp.take<int>(nn); a.parent_leaf_tmp = tmp.take<int>(nn);
    a.need2 = tmp.take<int>(ni);
    a.need4 = tmp.take<int>(ni);
    int* ghist = tmp.take<int>(radix_hist_ints(n));
    if (!tmp.base) return YRT_OK;   // dry pass

    a.nodes2 = io.nodes2; a.nodes4 = io.nodes4; a.ref_offset = io.ref_offset; a.size_bits = io.size_bits; a.rotate_pairs = io.rotate_pairs;
    a.seg_root = io.seg_root; a.seg_depth = io.seg_depth; a.seg_need2 = io.seg_need2; a.seg_need4 = io.seg_need4; a.leaf_size = io.leaf_size;
    YRT_CUDA(cudaMemsetAsync(a.parent_leaf, 0xff, sizeof(int) * nn, st));   // -1: no parent
    YRT_CUDA(cudaMemsetAsync(io.nodes2, 0, sizeof(float4) * YRT_NODE_STRIDE(2) * ni, st));
    YRT_CUDA(cudaMemsetAsync(io.nodes4, 0, sizeof(float4) * YRT_NODE_STRIDE(4) * ni, st));
    k_seg_init<<<grid_for(n_seg), 256, 0, st>>>(a);
    if (n > 0) {
        k_seg_bounds<<<grid_for(n), 256, 0, st>>>(a);
        k_morton<<<grid_for(n), 256, 0, st>>>(a);
        int* order0 = a.order;
        YRT_TRY(radix_sort_pairs(st, a.keys, a.order, keys_alt, order_alt, n, ghist));
        // an even number of passes leaves the result in the original buffers
        if (a.order != order0) { set_error("radix sort ended in the wrong buffer"); return YRT_ERR_CUDA; }
        if (n > 1) {
            k_karras<<<grid_for(n - 1), 256, 0, st>>>(a);
            k_refit<<<grid_for(n), 256, 0, st>>>(a);
            for (int r = 0; r < io.rotate_rounds; r++) {
                YRT_CUDA(cudaMemsetAsync(a.flags, 0, sizeof(int) * ni, st));
                k_rotate_refit<<<grid_for(n), 256, 0, st>>>(a);
            }
            if (io.rotate_rounds > 0) {   // leaf slots in tree order again (leaf references are (first, count) ranges)
                k_relayout_slot<<<grid_for(n), 256, 0, st>>>(a);
                k_relayout_move<<<grid_for(n), 256, 0, st>>>(a);
                k_relayout_refs<<<grid_for(n - 1), 256, 0, st>>>(a);
                k_relayout_copy<<<grid_for(n), 256, 0, st>>>(a);
            }
            k_emit<<<grid_for(n - 1), 256, 0, st>>>(a);
        }
        k_single_root<<<grid_for(n_seg), 256, 0, st>>>(a);
        if (n > 1) {
            k_depth<<<grid_for(n), 256, 0, st>>>(a);
            YRT_CUDA(cudaMemsetAsync(a.flags, 0, sizeof(int) * ni, st));
            k_stackneed<<<grid_for(n), 256, 0, st>>>(a);
        }
    }
    YRT_CUDA(cudaGetLastError());
    return YRT_OK;
}

// rank tables into leaf order (the ranks are computed on a host thread while the trees are built)
__global__ void k_permute_prim_rank(int n, const int* __restrict__ order, const int* __restrict__ rank_in, int* __restrict__ rank_out) {
    int k = YRT_TID();
    if (k < n) rank_out[k] = rank_in[order[k]];
}
// slot_of[instance] = its slot in inst_recs (scenes traced through the reference's instance tree, RefTlas in yrt_scene.cuh)
__global__ void k_slot_of_inst(int n_active, const int* __restrict__ order, const int* __restrict__ active_inst, int* __restrict__ slot_of) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_active) return;
    slot_of[active_inst[order[k]]] = k;
}

__global__ void k_permute_inst_rank(int n, const int* __restrict__ order, const int* __restrict__ active_inst, const int* __restrict__ rank_in,
                                    int* __restrict__ rank_out) {
    int k = YRT_TID();
    if (k < n) rank_out[k] = rank_in[active_inst[order[k]]];
}

// ------------------------------------------------------------------------------------------
// upload + build on one device
// ------------------------------------------------------------------------------------------
// The build arena (geometry inputs + LBVH temporaries) is kept per device for the life of the process and reused by the next
// scene: cudaFree synchronises the device and took 1 - 900 ms here, far more than the build itself.
static std::mutex g_build_mu;
static DevBuf g_build_arena[64];

// Scene arenas of destroyed scenes, kept per device for the next scene of a similar size: a process that rebuilds its scene
// (moving geometry) would otherwise pay one cudaFree and one cudaMalloc of tens of MB per build (measured here: 0.6 - 40 ms
// and 1 - 900 ms).  At most three blocks of at most 512 MB per device are kept; everything else is freed as before.
static std::mutex g_pool_mu;
static std::vector<std::pair<void*, size_t>> g_arena_pool[64];

static int arena_take(DevBuf& b, size_t n, int dev) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        auto& pool = g_arena_pool[dev];
        for (size_t i = 0; i < pool.size(); i++)
            if (pool[i].second >= n && pool[i].second <= 2 * n + (1u << 20)) {
                b.release();
                b.p = pool[i].first; b.bytes = pool[i].second; b.device = dev;
                pool.erase(pool.begin() + i);
                return YRT_OK;
            }
    }
    return b.alloc(n, dev);
}
void arena_give_back(DevBuf& b) {
    if (!b.p) return;
    if (b.device >= 0 && b.device < 64 && b.bytes <= (512u << 20)) {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        auto& pool = g_arena_pool[b.device];
        if (pool.size() < 3) {
            pool.push_back({b.p, b.bytes});
            b.p = nullptr; b.bytes = 0;
            return;
        }
    }
    b.release();
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int build_device_scene(HostScene& hs, int device, DevScene& ds) {
    if (device < 0 || device >= 64) { set_error("device index out of range"); return YRT_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(g_build_mu);
    DevBuf& build_arena = g_build_arena[device];
    const bool trace = getenv("YRT_BUILD_TRACE") != nullptr;
    double t0 = now_ms();
    ds.device = device;
    YRT_CUDA(cudaSetDevice(device));
    YRT_CUDA(cudaDeviceGetAttribute(&ds.sm_count, cudaDevAttrMultiProcessorCount, device));
    if (!ds.stream) YRT_CUDA(cudaStreamCreateWithFlags(&ds.stream, cudaStreamNonBlocking));
    cudaStream_t st = ds.stream;

    auto env_or = [](const char* name, int def, int lo, int hi) { const char* e = getenv(name); return std::min(std::max(e ? atoi(e) : def, lo), hi); };
    const int leaf_blas = env_or("YRT_LEAF_BLAS", YRT_LEAF_SIZE_BLAS, 1, YRT_LEAF_MAX_COUNT), leaf_tlas = env_or("YRT_LEAF_TLAS", YRT_LEAF_SIZE_TLAS, 1, YRT_LEAF_MAX_COUNT);
    const int size_bits_blas = env_or("YRT_SIZE_BITS_BLAS", YRT_SIZE_BITS_BLAS, 0, 3), size_bits_tlas = env_or("YRT_SIZE_BITS_TLAS", YRT_SIZE_BITS_TLAS, 0, 3);
    const int rotate_blas = env_or("YRT_ROTATE_BLAS", YRT_ROTATE_ROUNDS_BLAS, 0, 8), rotate_tlas = env_or("YRT_ROTATE_TLAS", YRT_ROTATE_ROUNDS_TLAS, 0, 8);
    const int rotate_pairs_blas = env_or("YRT_ROTATE_PAIRS_BLAS", YRT_ROTATE_PAIRS_BLAS, 0, 1), rotate_pairs_tlas = env_or("YRT_ROTATE_PAIRS_TLAS", YRT_ROTATE_PAIRS_TLAS, 0, 1);

    const int np = hs.n_prims, na = (int)hs.active_inst.size(), nsh = std::max(hs.n_shapes, 1);
    const int nb_int = np > 1 ? np - 1 : 0, nt_int = na > 1 ? na - 1 : 0;
    ds.n_prims = np; ds.n_active = na; ds.has_reflective = hs.n_reflective > 0;

    // ---- layout: persistent arena `keep`, build arena `tmp`; the same code runs dry (sizes) and real (pointers) ----
    struct Up { void* dst; const void* src; size_t bytes; };
    std::vector<Up> ups;
    GeomView g;
    SceneView& v = ds.view;
    int *d_inst_shape = nullptr, *d_inst_mat = nullptr, *d_active = nullptr, *d_prim_rank_in = nullptr, *d_inst_rank_in = nullptr, *d_has_uv = nullptr,
        *d_tl_seg_of = nullptr, *d_tl_seg_first = nullptr, *d_results = nullptr, *d_shape_kind = nullptr;
    float* d_inst_frame = nullptr;
    float4 *plo = nullptr, *phi = nullptr, *ilo = nullptr, *ihi = nullptr, *d_nodes2 = nullptr, *d_nodes4 = nullptr, *d_prim_recs = nullptr, *d_prim_attrs = nullptr, *d_inst_recs = nullptr;
    int *d_prim_rank = nullptr, *d_inst_rank = nullptr;
    float4* d_inst_box = nullptr;
    // scenes with non-rigid instance frames: room for the reference's instance tree (at most 2 n - 1 nodes, scene.cpp:646), filled in
    // once the rank thread has built it
    float4* d_ref_nodes = nullptr;
    int *d_ref_leaf = nullptr, *d_slot_of = nullptr;
    const bool ref_tlas = !hs.all_rigid && hs.n_instances > 0;
    // apex grids of the point lights (yrt_pgrid.cuh).  A light converges its shadow rays on ONE point only if its frame does
    // not rotate: shade() aims at transform_point(frame, pos - p) (src/raytrace.cpp:129-130), i.e. at (pos + o) - p for an
    // identity rotation; other lights, lights beyond the first YRT_MAX_LIGHT_GRIDS and scenes with non-rigid instance
    // frames keep the instance tree.
    const int n_lights = (int)hs.light_inst.size();
    // (an instance tree of a few dozen leaves is 5 levels deep: nothing to save, and the per-frame camera grid would cost more than it gains;
    //  YRT_PGRID_MIN_INSTANCES moves the threshold, the tests set it to 1 to exercise the grids on small scenes)
    ds.grids_allowed = hs.all_rigid && env_or("YRT_PGRID", 1, 0, 1) != 0 && na >= env_or("YRT_PGRID_MIN_INSTANCES", 32, 1, 1 << 30);
    ds.extent = hs.extent;
    int light_R = env_or("YRT_LIGHT_GRID_R", -1, -1, 1024);
    if (light_R < 0) { light_R = 8; while (light_R < 256 && light_R * light_R < na) light_R *= 2; }   // ~ one cell per instance and face
    if (!ds.grids_allowed) light_R = 0;
    ds.light_grid_R = light_R;
    PGridArrays lga[YRT_MAX_LIGHT_GRIDS];
    bool lg_on[YRT_MAX_LIGHT_GRIDS];
    int n_light_grids = 0;
    for (int k = 0; k < YRT_MAX_LIGHT_GRIDS; k++) {
        ds.light_grids.g[k] = gridref_none();
        lg_on[k] = false;
        if (k >= n_lights || light_R <= 0) continue;
        const float4* lr = &hs.light_recs[5 * (size_t)k];
        const bool identity = lr[0].x == 1.f && lr[0].y == 0.f && lr[0].z == 0.f && lr[1].x == 0.f && lr[1].y == 1.f && lr[1].z == 0.f &&
                              lr[2].x == 0.f && lr[2].y == 0.f && lr[2].z == 1.f;
        if (!identity) continue;
        const vec3 apex = mk3(lr[4].x + lr[3].x, lr[4].y + lr[3].y, lr[4].z + lr[3].z);
        const float reach = hs.extent + fmaxf(fmaxf(fabsf(apex.x), fabsf(apex.y)), fabsf(apex.z));
        if (!(reach < 1.0e18f)) continue;
        lga[k].d = pgrid_cube_desc(apex, light_R, reach, env_or("YRT_LIGHT_GRID_KEYS", 4 * 6 * light_R * light_R, 1, 1 << 28));   // (the switch: tests of the overflow path)
        lg_on[k] = true;
        n_light_grids++;
    }
    // the chain nodes of the light grids live behind the two trees in the node array the any-hit rays walk
    const int grid_node_first = nb_int + nt_int + 2;
    // (cells that find no room — keys or chain nodes — start at the tree's root like cells that are not served; YRT_LIGHT_GRID_NODES: tests)
    const int light_node_cap = n_light_grids ? env_or("YRT_LIGHT_GRID_NODES", (int)std::min<long long>((long long)n_light_grids * 6 * light_R * light_R * 3 / 4 + 1024, 1 << 26), 1, 1 << 26) : 0;
    const size_t n_nodes4 = (size_t)grid_node_first + (YRT_WIDE_ANY == 4 ? light_node_cap : 0);
    const size_t n_nodes2 = (size_t)grid_node_first + (YRT_WIDE_ANY == 2 ? light_node_cap : 0);
    int* d_light_node_total = nullptr;
    LbvhIo bo, to;
    const int tl_sf[2] = {0, na};
    // results block read back in one copy: [blas depth | blas need (binary) | blas need (4-wide) | tlas depth, need2, need4, root]
    const size_t n_results = 3 * (size_t)nsh + 4;
    Arena keep, tmp;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            YRT_TRY(arena_take(ds.arena, keep.off + 256, device));
            YRT_TRY(build_arena.alloc(tmp.off + 256, device));
            keep.base = (char*)ds.arena.p; tmp.base = (char*)build_arena.p;
            keep.off = tmp.off = 0;
        }
        ups.clear();
        auto up_keep = [&](auto* tag, const auto& vec) {
            using T = typename std::remove_pointer<decltype(tag)>::type;
            T* p = keep.take<T>(vec.size());
            ups.push_back({p, vec.data(), sizeof(T) * vec.size()});
            return p;
        };
        auto up_tmp = [&](auto* tag, const auto& vec) {
            using T = typename std::remove_pointer<decltype(tag)>::type;
            T* p = tmp.take<T>(vec.size());
            ups.push_back({p, vec.data(), sizeof(T) * vec.size()});
            return p;
        };
        // products and what the kernels of a frame read
        d_nodes4 = keep.take<float4>(YRT_NODE_STRIDE(4) * n_nodes4);   // (first in the arena: 256-byte aligned, one line per record)
        d_nodes2 = keep.take<float4>(YRT_NODE_STRIDE(2) * n_nodes2);
        d_prim_recs = keep.take<float4>(3 * (size_t)std::max(np, 1));
        d_prim_attrs = keep.take<float4>(YRT_ATTR_STRIDE * (size_t)std::max(np, 1));
        d_inst_recs = keep.take<float4>(4 * (size_t)std::max(na, 1));
        d_prim_rank = keep.take<int>((size_t)std::max(np, 1));
        d_inst_rank = keep.take<int>((size_t)std::max(na, 1));
        d_inst_box = keep.take<float4>(2 * (size_t)std::max(na, 1));
        if (ref_tlas) {
            d_ref_nodes = keep.take<float4>(2 * (size_t)(2 * hs.n_instances));
            d_ref_leaf = keep.take<int>((size_t)hs.n_instances);
            d_slot_of = keep.take<int>((size_t)hs.n_instances);
        }
        for (int k = 0; k < YRT_MAX_LIGHT_GRIDS; k++) {
            if (!lg_on[k]) continue;
            PGridArrays& a = lga[k];
            a.roots = keep.take<int>((size_t)a.d.n_cells);
            a.cells = tmp.take<int2>((size_t)a.d.n_cells);
            a.cnt = tmp.take<int>((size_t)a.d.n_cells + 2 + (size_t)na);
            a.total = a.cnt ? a.cnt + a.d.n_cells : nullptr;
            a.big = a.cnt ? a.cnt + a.d.n_cells + 1 : nullptr;
            a.keys = tmp.take<unsigned long long>((size_t)a.d.capacity);
            a.inst_box = d_inst_box;
            a.n_inst = na;
            a.as_list = 0; a.fixed_k = 0; a.entries = nullptr;
            a.arity = YRT_WIDE_ANY; a.nodes = YRT_WIDE_ANY == 4 ? d_nodes4 : d_nodes2;
            a.node_first = grid_node_first; a.node_capacity = light_node_cap;
        }
        d_light_node_total = tmp.take<int>(1);
        v.mat_recs = up_keep((float4*)nullptr, hs.mat_recs);
        v.light_recs = up_keep((float4*)nullptr, hs.light_recs);
        v.tex_rgba8 = up_keep((uint8_t*)nullptr, hs.tex_rgba8);
        v.tex_info = up_keep((int4*)nullptr, hs.tex_info);
        {
            float* lut = keep.take<float>(256);
            ups.push_back({lut, hs.srgb_lut, sizeof(hs.srgb_lut)});
            v.srgb_lut = lut;
        }
        d_results = keep.take<int>(n_results);
        // geometry inputs: only the build reads them
        d_shape_kind = up_tmp((int*)nullptr, hs.shape_kind);
        g.shape_kind = d_shape_kind;
        g.shape_elem_off = up_tmp((int*)nullptr, hs.shape_elem_off);
        g.shape_elem_cnt = up_tmp((int*)nullptr, hs.shape_elem_cnt);
        g.shape_vert_off = up_tmp((int*)nullptr, hs.shape_vert_off);
        g.shape_prim_off = up_tmp((int*)nullptr, hs.shape_prim_off);
        g.elem_idx = up_tmp((int*)nullptr, hs.elem_idx);
        g.pos = up_tmp((float*)nullptr, hs.pos);
        g.norm = up_tmp((float*)nullptr, hs.norm);
        g.uv = up_tmp((float*)nullptr, hs.uv);
        g.radius = up_tmp((float*)nullptr, hs.radius);
        g.prim_shape = up_tmp((int*)nullptr, hs.prim_shape);
        g.n_prims = np;
        d_inst_frame = up_tmp((float*)nullptr, hs.inst_frame);
        d_inst_shape = up_tmp((int*)nullptr, hs.inst_shape);
        d_inst_mat = up_tmp((int*)nullptr, hs.inst_mat);
        d_active = up_tmp((int*)nullptr, hs.active_inst);
        d_has_uv = up_tmp((int*)nullptr, hs.shape_has_uv);
        d_prim_rank_in = tmp.take<int>((size_t)std::max(np, 1));        // uploaded late (rank thread)
        d_inst_rank_in = tmp.take<int>((size_t)std::max(hs.n_instances, 1));
        d_tl_seg_of = tmp.take<int>((size_t)std::max(na, 1));
        d_tl_seg_first = tmp.take<int>(2);
        ups.push_back({d_tl_seg_first, tl_sf, sizeof(tl_sf)});
        plo = tmp.take<float4>((size_t)std::max(np, 1)); phi = tmp.take<float4>((size_t)std::max(np, 1));
        ilo = tmp.take<float4>((size_t)std::max(na, 1)); ihi = tmp.take<float4>((size_t)std::max(na, 1));
        int* shape_box_lo = tmp.take<int>(3 * (size_t)nsh);
        int* shape_box_hi = tmp.take<int>(3 * (size_t)nsh);
        int* blas_root = tmp.take<int>((size_t)nsh);
        int* tl_box_lo = tmp.take<int>(3);
        int* tl_box_hi = tmp.take<int>(3);
        // ---- BLAS: all shapes in one build ----
        bo.n = np; bo.n_seg = nsh; bo.leaf_size = leaf_blas; bo.ref_offset = 0; bo.size_bits = size_bits_blas; bo.rotate_rounds = rotate_blas; bo.rotate_pairs = rotate_pairs_blas;
        bo.box_lo = plo; bo.box_hi = phi; bo.seg_of = g.prim_shape; bo.seg_first = g.shape_prim_off; bo.nodes2 = d_nodes2; bo.nodes4 = d_nodes4;
        bo.seg_root = blas_root; bo.seg_depth = d_results; bo.seg_need2 = d_results + nsh; bo.seg_need4 = d_results + 2 * nsh; bo.seg_box_lo = shape_box_lo; bo.seg_box_hi = shape_box_hi;
        // ---- TLAS over active instances ----
        to.n = na; to.n_seg = 1; to.leaf_size = leaf_tlas; to.ref_offset = nb_int; to.size_bits = size_bits_tlas; to.rotate_rounds = rotate_tlas; to.rotate_pairs = rotate_pairs_tlas;
        to.box_lo = ilo; to.box_hi = ihi; to.seg_of = d_tl_seg_of; to.seg_first = d_tl_seg_first; to.nodes2 = d_nodes2 ? d_nodes2 + YRT_NODE_STRIDE(2) * (size_t)nb_int : nullptr; to.nodes4 = d_nodes4 ? d_nodes4 + YRT_NODE_STRIDE(4) * (size_t)nb_int : nullptr;
        to.seg_root = d_results + 3 * nsh + 3; to.seg_depth = d_results + 3 * nsh; to.seg_need2 = d_results + 3 * nsh + 1; to.seg_need4 = d_results + 3 * nsh + 2; to.seg_box_lo = tl_box_lo; to.seg_box_hi = tl_box_hi;
        if (pass == 0) {   // sizes of the two builds' temporaries (not shared: the BLAS order is read again at the very end)
            YRT_TRY(lbvh_build(st, tmp, bo));
            YRT_TRY(lbvh_build(st, tmp, to));
        }
    }
    const double t_layout = now_ms();
    for (const Up& u : ups)
        if (u.bytes) YRT_CUDA(cudaMemcpyAsync(u.dst, u.src, u.bytes, cudaMemcpyHostToDevice, st));
    YRT_CUDA(cudaMemsetAsync(d_tl_seg_of, 0, sizeof(int) * (size_t)std::max(na, 1), st));
    const double t_upload = now_ms();

    cudaEvent_t e0, e1;
    YRT_CUDA(cudaEventCreate(&e0));
    YRT_CUDA(cudaEventCreate(&e1));
    YRT_CUDA(cudaEventRecord(e0, st));
    if (np > 0) k_prim_boxes<<<grid_for(np), 256, 0, st>>>(g, plo, phi);
    YRT_TRY(lbvh_build(st, tmp, bo));
    if (np > 0) k_gather_prims<<<grid_for(np), 256, 0, st>>>(g, bo.order, d_prim_recs, d_prim_attrs, d_has_uv);
    if (na > 0) k_inst_boxes<<<grid_for(na), 256, 0, st>>>(na, d_active, d_inst_frame, d_inst_shape, bo.seg_box_lo, bo.seg_box_hi, ilo, ihi);
    YRT_TRY(lbvh_build(st, tmp, to));
    if (na > 0) k_inst_recs<<<grid_for(na), 256, 0, st>>>(na, to.order, d_active, d_inst_frame, d_inst_shape, d_inst_mat, d_shape_kind, bo.seg_root, d_inst_recs);
    if (na > 0) k_inst_box<<<grid_for(na), 256, 0, st>>>(na, to.order, ilo, ihi, d_inst_box);
    if (n_light_grids) YRT_CUDA(cudaMemsetAsync(d_light_node_total, 0, sizeof(int), st));
    for (int k = 0; k < YRT_MAX_LIGHT_GRIDS; k++) {
        if (!lg_on[k]) continue;
        lga[k].node_total = d_light_node_total;
        lga[k].tlas_root = 0;
        lga[k].tlas_root_dev = to.seg_root;      // (the root is only known on the device at this point)
        YRT_TRY(pgrid_build_enqueue(lga[k], st));
        GridRef& g = ds.light_grids.g[k];
        g = gridref_none();
        g.roots = lga[k].roots; g.nx = light_R;
    }
    YRT_CUDA(cudaGetLastError());
    const double t_enqueued = now_ms();
    // tie-break ranks: computed on a host thread since yrt_scene_create started (yrt_host.cu); needed only now
    hs.wait_ranks();
    const double t_ranks = now_ms();
    YRT_CUDA(cudaMemcpyAsync(d_prim_rank_in, hs.prim_rank.data(), sizeof(int) * hs.prim_rank.size(), cudaMemcpyHostToDevice, st));
    YRT_CUDA(cudaMemcpyAsync(d_inst_rank_in, hs.inst_rank.data(), sizeof(int) * hs.inst_rank.size(), cudaMemcpyHostToDevice, st));
    if (np > 0) k_permute_prim_rank<<<grid_for(np), 256, 0, st>>>(np, bo.order, d_prim_rank_in, d_prim_rank);
    if (na > 0) k_permute_inst_rank<<<grid_for(na), 256, 0, st>>>(na, to.order, d_active, d_inst_rank_in, d_inst_rank);
    if (ref_tlas) {
        if (hs.ref_nodes.size() > 2 * (size_t)(2 * hs.n_instances) || hs.ref_leaf_inst.size() != (size_t)hs.n_instances) { set_error("internal: reference instance tree out of bounds"); return YRT_ERR_CUDA; }
        YRT_CUDA(cudaMemcpyAsync(d_ref_nodes, hs.ref_nodes.data(), sizeof(float4) * hs.ref_nodes.size(), cudaMemcpyHostToDevice, st));
        YRT_CUDA(cudaMemcpyAsync(d_ref_leaf, hs.ref_leaf_inst.data(), sizeof(int) * hs.ref_leaf_inst.size(), cudaMemcpyHostToDevice, st));
        YRT_CUDA(cudaMemsetAsync(d_slot_of, 0xff, sizeof(int) * (size_t)hs.n_instances, st));
        if (na > 0) k_slot_of_inst<<<grid_for(na), 256, 0, st>>>(na, to.order, d_active, d_slot_of);
    }
    YRT_CUDA(cudaEventRecord(e1, st));
    std::vector<int> results(n_results, 0);
    YRT_CUDA(cudaMemcpyAsync(results.data(), d_results, sizeof(int) * n_results, cudaMemcpyDeviceToHost, st));
    YRT_CUDA(cudaStreamSynchronize(st));     // the only wait of the build
    const double t_done = now_ms();
    float ms = 0.f;
    YRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    ds.build_us = ms * 1000.f;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (trace)
        fprintf(stderr, "[yrt build] device %d: layout %.3f ms, uploads enqueued %.3f, build enqueued %.3f, waited for ranks %.3f, device done %.3f (kernels %.3f); "
                        "arenas %.1f + %.1f MB\n", device, t_layout - t0, t_upload - t_layout, t_enqueued - t_upload, t_ranks - t_enqueued, t_done - t_ranks, ms,
                keep.off / 1e6, tmp.off / 1e6);

    ds.blas_depth = *std::max_element(results.begin(), results.begin() + nsh);
    const int blas_need2 = *std::max_element(results.begin() + nsh, results.begin() + 2 * nsh);
    const int blas_need4 = *std::max_element(results.begin() + 2 * nsh, results.begin() + 3 * nsh);
    ds.tlas_depth = results[3 * nsh];
    const int tlas_need2 = results[3 * nsh + 1], tlas_need4 = results[3 * nsh + 2], troot = results[3 * nsh + 3];
    ds.n_blas_nodes = nb_int;
    ds.n_tlas_nodes = nt_int;
    if (na > 0 && troot == YRT_REF_SENTINEL) { set_error("internal: TLAS root not found"); return YRT_ERR_CUDA; }
    // entries below the TLAS root + entries below the deepest BLAS root + guard + sentinel + the "rest of a TLAS leaf" entry
    const int need2 = tlas_need2 + blas_need2 + 3, need4 = tlas_need4 + blas_need4 + 3;
    ds.stack_need = std::max(YRT_WIDE_CLOSEST == 4 ? need4 : need2, YRT_WIDE_ANY == 4 ? need4 : need2);
    if (ds.stack_need > YRT_STACK_CAP) {
        set_error("traversal tree too deep for the stack (%d entries needed: tlas %d + blas %d levels, capacity %d)", ds.stack_need, ds.tlas_depth, ds.blas_depth, YRT_STACK_CAP);
        return YRT_ERR_UNSUPPORTED;
    }
    // a cell's chain postpones up to YRT_PGRID_MAX_LIST + 2 entries on a ray's stack where the instance tree postpones its depth's worth
    {
        const int chain = YRT_PGRID_MAX_LIST + 2;
        const int g2 = std::max(tlas_need2, chain) + blas_need2 + 3, g4 = std::max(tlas_need4, chain) + blas_need4 + 3;
        if (ds.grids_allowed && std::max(YRT_WIDE_CLOSEST == 4 ? g4 : g2, YRT_WIDE_ANY == 4 ? g4 : g2) > YRT_STACK_CAP) {
            ds.grids_allowed = false;
            for (int k = 0; k < YRT_MAX_LIGHT_GRIDS; k++) ds.light_grids.g[k] = gridref_none();
        }
    }
    v.nodes2 = d_nodes2;
    v.nodes4 = d_nodes4;
    v.inst_recs = d_inst_recs;
    v.inst_box = d_inst_box;
    v.prim_recs = d_prim_recs;
    v.prim_attrs = d_prim_attrs;
    v.inst_rank = d_inst_rank;
    v.prim_rank = d_prim_rank;
    v.tlas_root = troot;
    v.n_lights = (int)hs.light_inst.size();
    v.n_active_instances = na;
    ds.ref.nodes = d_ref_nodes;
    ds.ref.leaf_inst = d_ref_leaf;
    ds.ref.slot_of_inst = d_slot_of;
    ds.ref.n_nodes = ref_tlas ? (int)(hs.ref_nodes.size() / 2) : 0;
    return YRT_OK;
}

// The apex grid of one frame's camera (pinhole = eval_camera's ray origin, src/raytrace.cpp:6-37): cells are tiles of
// 2^shift x 2^shift pixels.  Rebuilt for every frame — the camera is an argument of the render call — on `st` (the scene's grid
// stream, see render_rows_device): a memset and three small launches, no host synchronisation.
int camera_grid_enqueue(DevScene& ds, const camera_k& cam, int width, int height, cudaStream_t st, GridRef* out) {
    *out = gridref_none();
    auto env_or = [](const char* name, int def, int lo, int hi) { const char* e = getenv(name); return std::min(std::max(e ? atoi(e) : def, lo), hi); };
    if (!ds.grids_allowed || ds.n_active <= 0 || env_or("YRT_CAM_GRID", 1, 0, 1) == 0) return YRT_OK;
    int shift = env_or("YRT_CAM_CELL_SHIFT", 3, 0, 8);
    const float reach = ds.extent + fmaxf(fmaxf(fabsf(cam.frame.o.x), fabsf(cam.frame.o.y)), fabsf(cam.frame.o.z));
    if (!(reach < 1.0e18f)) return YRT_OK;
    auto cells_at = [&](int sh) { return (long long)((width + (1 << sh) - 1) >> sh) * ((height + (1 << sh) - 1) >> sh); };
    while (shift < 12 && cells_at(shift) > (1ll << 18)) shift++;      // at most 262 144 cells (8x8-pixel cells up to 4K frames): 16 slots x 40 B each
    const long long cells = cells_at(shift);
    const int K = 16;                                                 // slots per cell (YRT_PGRID_MAX_LIST or fewer candidates are served)
    const int capacity = (int)(K * cells);
    PGridArrays a;
    a.d = pgrid_camera_desc(cam, width, height, shift, reach, capacity);
    a.d.max_list = std::min(a.d.max_list, K);
    if (a.d.mode == 0) return YRT_OK;
    YRT_TRY(ds.cg_cells.alloc(sizeof(int2) * (size_t)a.d.n_cells, ds.device));
    YRT_TRY(ds.cg_entries.alloc(sizeof(float4) * 2 * (size_t)capacity, ds.device));
    YRT_TRY(ds.cg_keys.alloc(sizeof(unsigned long long) * (size_t)capacity, ds.device));
    YRT_TRY(ds.cg_cnt.alloc(sizeof(int) * ((size_t)a.d.n_cells + 2 + (size_t)ds.n_active), ds.device));
    a.inst_box = ds.view.inst_box; a.n_inst = ds.n_active;
    a.cnt = ds.cg_cnt.as<int>(); a.total = a.cnt + a.d.n_cells; a.big = a.cnt + a.d.n_cells + 1;
    a.cells = ds.cg_cells.as<int2>(); a.keys = ds.cg_keys.as<unsigned long long>();
    a.as_list = 1; a.fixed_k = K; a.entries = ds.cg_entries.as<float4>();
    a.roots = nullptr; a.nodes = nullptr; a.arity = 0; a.node_first = a.node_capacity = 0; a.node_total = nullptr;
    a.tlas_root = ds.view.tlas_root; a.tlas_root_dev = nullptr;
    YRT_TRY(pgrid_build_enqueue(a, st));
    out->cells = a.cells; out->entries = a.entries; out->nx = a.d.nx; out->shift = shift;
    return YRT_OK;
}

}  // namespace yrt
