// yrt_build.cu — scene upload and the GPU two-level LBVH build (sm_100a).
//
// Replaces build_bvh(scn,false) (src/raytrace.cpp:278 -> src/scene.cpp:554-565 and :525-549,
// :572-658).  Launch sequence per tree set (BLAS over all shapes' elements at once, then TLAS
// over instances): bounds -> per-segment reductions -> Morton keys -> 8 x {histogram, scan,
// stable scatter} LSD radix sort -> Karras topology -> atomic bottom-up refit -> node emit.
#include <algorithm>
#include <chrono>
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
                               float4* __restrict__ prim_attrs, const int* __restrict__ shape_has_uv,
                               const int* __restrict__ rank_in, int* __restrict__ rank_out) {
    int k = YRT_TID();
    if (k >= g.n_prims) return;
    int gp = order[k];
    rank_out[k] = rank_in[gp];
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
                            const int* __restrict__ blas_root, float4* __restrict__ recs,
                            const int* __restrict__ rank_in, int* __restrict__ rank_out) {
    int k = YRT_TID();
    if (k >= n_active) return;
    int inst = active_inst[order[k]];
    rank_out[k] = rank_in[inst];
    int s = inst_shape[inst];
    const float* fr = inst_frame + 12 * (size_t)inst;
    float4* r = recs + 4 * (size_t)k;
    r[0] = mk4(fr[0], fr[1], fr[2], int_as_float(blas_root[s]));
    r[1] = mk4(fr[3], fr[4], fr[5], int_as_float(inst));
    r[2] = mk4(fr[6], fr[7], fr[8], int_as_float(inst_mat[inst]));
    r[3] = mk4(fr[9], fr[10], fr[11], int_as_float((int)((unsigned)s | ((unsigned)shape_kind[s] << 28))));
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

static int radix_sort_pairs(int dev, cudaStream_t st, unsigned long long*& keys, int*& vals, unsigned long long*& keys_alt,
                            int*& vals_alt, int n, DevBuf& ghist_buf) {
    if (n <= 1) return YRT_OK;
    int nblocks = (n + RS_TILE - 1) / RS_TILE;
    YRT_TRY(ghist_buf.alloc(sizeof(int) * 256 * (size_t)nblocks, dev));
    int* ghist = ghist_buf.as<int>();
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

// exported for tests (python ctypes): sort n (key, value) pairs held in HOST arrays on the device
extern "C" int yrt_debug_sort_pairs(unsigned long long* h_keys, int* h_vals, int n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return YRT_ERR_NO_DEVICE; }
    DevBuf k0, k1, v0, v1, gh;
    YRT_TRY(k0.upload(h_keys, sizeof(unsigned long long) * (size_t)n, dev, 0));
    YRT_TRY(v0.upload(h_vals, sizeof(int) * (size_t)n, dev, 0));
    YRT_TRY(k1.alloc(sizeof(unsigned long long) * (size_t)n, dev));
    YRT_TRY(v1.alloc(sizeof(int) * (size_t)n, dev));
    unsigned long long *ka = k0.as<unsigned long long>(), *kb = k1.as<unsigned long long>();
    int *va = v0.as<int>(), *vb = v1.as<int>();
    YRT_TRY(radix_sort_pairs(dev, 0, ka, va, kb, vb, n, gh));
    YRT_CUDA(cudaMemcpy(h_keys, ka, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToHost));
    YRT_CUDA(cudaMemcpy(h_vals, va, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    return YRT_OK;
}

// ------------------------------------------------------------------------------------------
// one LBVH build (many segments)
// ------------------------------------------------------------------------------------------
struct LbvhOut {
    float4* nodes;        // YRT_NODE_STRIDE float4 per internal node, preallocated by the caller
    int ref_offset;       // index of this tree set's node 0 in the shared node array
    int size_bits;        // size-class bits in the sort key (see morton_item)
    int rotate_rounds;    // bottom-up tree-rotation passes after the refit (see rotate_refit_item)
    int rotate_pairs;     // ... that also try grandchild pair exchanges
    DevBuf* seg_root;     // [n_seg]
    DevBuf* seg_depth;    // [n_seg]
    DevBuf* seg_need;     // [n_seg] traversal stack entries needed below the segment root
    DevBuf* order;        // [n] item id at sorted slot
    DevBuf* seg_box_lo;   // [3*n_seg] ordered ints (may be null)
    DevBuf* seg_box_hi;
};

static inline int grid_for(int n, int t = 256) { return n > 0 ? (n + t - 1) / t : 1; }

static int lbvh_build(int dev, cudaStream_t st, int n, int n_seg, float4* box_lo, float4* box_hi, const int* d_seg_of,
                      const int* d_seg_first, int leaf_size, LbvhOut& out) {
    DevBuf cent_lo, cent_hi, sbox_lo_tmp, sbox_hi_tmp, keys, keys_alt, order_alt, left, right, rfirst, rlast, pint, pleaf,
        flags, nlo, nhi, ghist, count, new_slot, order_tmp, pleaf_tmp, need;
    DevBuf* sbl = out.seg_box_lo ? out.seg_box_lo : &sbox_lo_tmp;
    DevBuf* sbh = out.seg_box_hi ? out.seg_box_hi : &sbox_hi_tmp;
    size_t ni = n > 1 ? (size_t)(n - 1) : 1;
    YRT_TRY(cent_lo.alloc(sizeof(int) * 3 * (size_t)n_seg, dev));
    YRT_TRY(cent_hi.alloc(sizeof(int) * 3 * (size_t)n_seg, dev));
    YRT_TRY(sbl->alloc(sizeof(int) * 3 * (size_t)n_seg, dev));
    YRT_TRY(sbh->alloc(sizeof(int) * 3 * (size_t)n_seg, dev));
    YRT_TRY(keys.alloc(sizeof(unsigned long long) * (size_t)std::max(n, 1), dev));
    YRT_TRY(keys_alt.alloc(sizeof(unsigned long long) * (size_t)std::max(n, 1), dev));
    YRT_TRY(out.order->alloc(sizeof(int) * (size_t)std::max(n, 1), dev));
    YRT_TRY(order_alt.alloc(sizeof(int) * (size_t)std::max(n, 1), dev));
    YRT_TRY(left.alloc(sizeof(int) * ni, dev));
    YRT_TRY(right.alloc(sizeof(int) * ni, dev));
    YRT_TRY(rfirst.alloc(sizeof(int) * ni, dev));
    YRT_TRY(rlast.alloc(sizeof(int) * ni, dev));
    YRT_TRY(pint.alloc(sizeof(int) * ni, dev));
    YRT_TRY(pleaf.alloc(sizeof(int) * (size_t)std::max(n, 1), dev));
    YRT_TRY(flags.alloc(sizeof(int) * ni, dev));
    YRT_TRY(nlo.alloc(sizeof(float4) * ni, dev));
    YRT_TRY(nhi.alloc(sizeof(float4) * ni, dev));
    if (out.rotate_rounds > 0) {
        YRT_TRY(count.alloc(sizeof(int) * ni, dev));
        YRT_TRY(new_slot.alloc(sizeof(int) * (size_t)std::max(n, 1), dev));
        YRT_TRY(order_tmp.alloc(sizeof(int) * (size_t)std::max(n, 1), dev));
        YRT_TRY(pleaf_tmp.alloc(sizeof(int) * (size_t)std::max(n, 1), dev));
    }
    YRT_TRY(out.seg_root->alloc(sizeof(int) * (size_t)n_seg, dev));
    YRT_TRY(out.seg_depth->alloc(sizeof(int) * (size_t)n_seg, dev));
    YRT_TRY(out.seg_need->alloc(sizeof(int) * (size_t)n_seg, dev));
    YRT_TRY(need.alloc(sizeof(int) * ni, dev));
    YRT_CUDA(cudaMemsetAsync(pleaf.p, 0xff, sizeof(int) * (size_t)std::max(n, 1), st));   // -1: no parent
    YRT_CUDA(cudaMemsetAsync(out.nodes, 0, sizeof(float4) * YRT_NODE_STRIDE * ni, st));

    LbvhArrays a;
    a.n = n;
    a.n_seg = n_seg;
    a.box_lo = box_lo;
    a.box_hi = box_hi;
    a.seg_of = d_seg_of;
    a.seg_first = d_seg_first;
    a.seg_cent_lo = cent_lo.as<int>();
    a.seg_cent_hi = cent_hi.as<int>();
    a.seg_box_lo = sbl->as<int>();
    a.seg_box_hi = sbh->as<int>();
    a.keys = keys.as<unsigned long long>();
    a.order = out.order->as<int>();
    a.left = left.as<int>();
    a.right = right.as<int>();
    a.range_first = rfirst.as<int>();
    a.range_last = rlast.as<int>();
    a.parent_int = pint.as<int>();
    a.parent_leaf = pleaf.as<int>();
    a.flags = flags.as<int>();
    a.count = count.as<int>();
    a.new_slot = new_slot.as<int>();
    a.order_tmp = order_tmp.as<int>();
    a.parent_leaf_tmp = pleaf_tmp.as<int>();
    a.node_lo = nlo.as<float4>();
    a.node_hi = nhi.as<float4>();
    a.nodes = out.nodes;
    a.ref_offset = out.ref_offset;
    a.size_bits = out.size_bits;
    a.rotate_pairs = out.rotate_pairs;
    a.seg_root = out.seg_root->as<int>();
    a.seg_depth = out.seg_depth->as<int>();
    a.seg_need = out.seg_need->as<int>();
    a.need = need.as<int>();
    a.leaf_size = leaf_size;

    k_seg_init<<<grid_for(n_seg), 256, 0, st>>>(a);
    if (n > 0) {
        k_seg_bounds<<<grid_for(n), 256, 0, st>>>(a);
        k_morton<<<grid_for(n), 256, 0, st>>>(a);
        unsigned long long* kalt = keys_alt.as<unsigned long long>();
        int* oalt = order_alt.as<int>();
        YRT_TRY(radix_sort_pairs(dev, st, a.keys, a.order, kalt, oalt, n, ghist));
        // an even number of passes leaves the result in the original buffers
        if (a.order != out.order->as<int>()) { set_error("radix sort ended in the wrong buffer"); return YRT_ERR_CUDA; }
        if (n > 1) {
            k_karras<<<grid_for(n - 1), 256, 0, st>>>(a);
            k_refit<<<grid_for(n), 256, 0, st>>>(a);
            for (int r = 0; r < out.rotate_rounds; r++) {
                YRT_CUDA(cudaMemsetAsync(a.flags, 0, sizeof(int) * ni, st));
                k_rotate_refit<<<grid_for(n), 256, 0, st>>>(a);
            }
            if (out.rotate_rounds > 0) {   // leaf slots in tree order again (leaf references are (first, count) ranges)
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
    // temporaries are freed when this returns: wait for the work that uses them
    YRT_CUDA(cudaStreamSynchronize(st));
    return YRT_OK;
}

// ------------------------------------------------------------------------------------------
// upload + build on one device
// ------------------------------------------------------------------------------------------
template <class T> static int up(DevBuf& b, const std::vector<T>& v, int dev, cudaStream_t st) {
    return b.upload(v.data(), sizeof(T) * v.size(), dev, st);
}

int build_device_scene(const HostScene& hs, int device, DevScene& ds) {
    ds.device = device;
    YRT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    YRT_CUDA(cudaGetDeviceProperties(&prop, device));
    ds.sm_count = prop.multiProcessorCount;
    if (!ds.stream) YRT_CUDA(cudaStreamCreateWithFlags(&ds.stream, cudaStreamNonBlocking));
    cudaStream_t st = ds.stream;

    YRT_TRY(up(ds.shape_kind, hs.shape_kind, device, st));
    YRT_TRY(up(ds.shape_elem_off, hs.shape_elem_off, device, st));
    YRT_TRY(up(ds.shape_elem_cnt, hs.shape_elem_cnt, device, st));
    YRT_TRY(up(ds.shape_vert_off, hs.shape_vert_off, device, st));
    YRT_TRY(up(ds.shape_prim_off, hs.shape_prim_off, device, st));
    YRT_TRY(up(ds.elem_idx, hs.elem_idx, device, st));
    YRT_TRY(up(ds.pos, hs.pos, device, st));
    YRT_TRY(up(ds.norm, hs.norm, device, st));
    YRT_TRY(up(ds.uv, hs.uv, device, st));
    YRT_TRY(up(ds.radius, hs.radius, device, st));
    YRT_TRY(up(ds.prim_shape, hs.prim_shape, device, st));
    YRT_TRY(up(ds.inst_frame, hs.inst_frame, device, st));
    YRT_TRY(up(ds.inst_shape, hs.inst_shape, device, st));
    YRT_TRY(up(ds.inst_mat, hs.inst_mat, device, st));
    YRT_TRY(up(ds.active_inst, hs.active_inst, device, st));
    YRT_TRY(up(ds.prim_rank_in, hs.prim_rank, device, st));
    YRT_TRY(up(ds.inst_rank_in, hs.inst_rank, device, st));
    YRT_TRY(up(ds.mat_recs, hs.mat_recs, device, st));
    YRT_TRY(up(ds.light_recs, hs.light_recs, device, st));
    YRT_TRY(up(ds.tex, hs.tex_rgba8, device, st));
    YRT_TRY(up(ds.tex_info, hs.tex_info, device, st));
    YRT_TRY(ds.lut.upload(hs.srgb_lut, sizeof(hs.srgb_lut), device, st));
    DevBuf shape_has_uv;
    YRT_TRY(up(shape_has_uv, hs.shape_has_uv, device, st));
    YRT_CUDA(cudaStreamSynchronize(st));   // uploads done: time the build alone

    cudaEvent_t e0, e1;
    YRT_CUDA(cudaEventCreate(&e0));
    YRT_CUDA(cudaEventCreate(&e1));
    YRT_CUDA(cudaEventRecord(e0, st));

    const char* env_lb = getenv("YRT_LEAF_BLAS");
    const char* env_lt = getenv("YRT_LEAF_TLAS");
    int leaf_blas = env_lb ? atoi(env_lb) : YRT_LEAF_SIZE_BLAS;
    int leaf_tlas = env_lt ? atoi(env_lt) : YRT_LEAF_SIZE_TLAS;
    const char* env_sb = getenv("YRT_SIZE_BITS_BLAS");
    const char* env_st = getenv("YRT_SIZE_BITS_TLAS");
    int size_bits_blas = std::min(std::max(env_sb ? atoi(env_sb) : YRT_SIZE_BITS_BLAS, 0), 3);
    int size_bits_tlas = std::min(std::max(env_st ? atoi(env_st) : YRT_SIZE_BITS_TLAS, 0), 3);
    const char* env_rb = getenv("YRT_ROTATE_BLAS");
    const char* env_rt = getenv("YRT_ROTATE_TLAS");
    int rotate_blas = std::min(std::max(env_rb ? atoi(env_rb) : YRT_ROTATE_ROUNDS_BLAS, 0), 8);
    int rotate_tlas = std::min(std::max(env_rt ? atoi(env_rt) : YRT_ROTATE_ROUNDS_TLAS, 0), 8);
    int rotate_pairs_blas = getenv("YRT_ROTATE_PAIRS_BLAS") ? atoi(getenv("YRT_ROTATE_PAIRS_BLAS")) : YRT_ROTATE_PAIRS_BLAS;
    int rotate_pairs_tlas = getenv("YRT_ROTATE_PAIRS_TLAS") ? atoi(getenv("YRT_ROTATE_PAIRS_TLAS")) : YRT_ROTATE_PAIRS_TLAS;
    leaf_blas = std::min(std::max(leaf_blas, 1), YRT_LEAF_MAX_COUNT);
    leaf_tlas = std::min(std::max(leaf_tlas, 1), YRT_LEAF_MAX_COUNT);

    GeomView g;
    g.shape_kind = ds.shape_kind.as<int>();
    g.shape_elem_off = ds.shape_elem_off.as<int>();
    g.shape_elem_cnt = ds.shape_elem_cnt.as<int>();
    g.shape_vert_off = ds.shape_vert_off.as<int>();
    g.shape_prim_off = ds.shape_prim_off.as<int>();
    g.elem_idx = ds.elem_idx.as<int>();
    g.pos = ds.pos.as<float>();
    g.norm = ds.norm.as<float>();
    g.uv = ds.uv.as<float>();
    g.radius = ds.radius.as<float>();
    g.prim_shape = ds.prim_shape.as<int>();
    g.n_prims = hs.n_prims;
    ds.n_prims = hs.n_prims;
    ds.n_active = (int)hs.active_inst.size();
    ds.has_reflective = hs.n_reflective > 0;

    // ---- BLAS: all shapes in one build ----
    DevBuf plo, phi, blas_order;
    YRT_TRY(plo.alloc(sizeof(float4) * (size_t)std::max(hs.n_prims, 1), device));
    YRT_TRY(phi.alloc(sizeof(float4) * (size_t)std::max(hs.n_prims, 1), device));
    if (hs.n_prims > 0) k_prim_boxes<<<grid_for(hs.n_prims), 256, 0, st>>>(g, plo.as<float4>(), phi.as<float4>());
    int nb_int = hs.n_prims > 1 ? hs.n_prims - 1 : 0, nt_int = ds.n_active > 1 ? ds.n_active - 1 : 0;
    YRT_TRY(ds.nodes.alloc(sizeof(float4) * YRT_NODE_STRIDE * (size_t)(nb_int + nt_int + 2), device));
    LbvhOut bo;
    bo.nodes = ds.nodes.as<float4>(); bo.ref_offset = 0; bo.size_bits = size_bits_blas; bo.rotate_rounds = rotate_blas; bo.rotate_pairs = rotate_pairs_blas; bo.seg_root = &ds.blas_seg_root; bo.seg_depth = &ds.blas_seg_depth; bo.seg_need = &ds.blas_seg_need; bo.order = &blas_order;
    bo.seg_box_lo = &ds.shape_box_lo; bo.seg_box_hi = &ds.shape_box_hi;
    YRT_TRY(lbvh_build(device, st, hs.n_prims, std::max(hs.n_shapes, 1), plo.as<float4>(), phi.as<float4>(), g.prim_shape,
                       g.shape_prim_off, leaf_blas, bo));
    YRT_TRY(ds.prim_recs.alloc(sizeof(float4) * 3 * (size_t)std::max(hs.n_prims, 1), device));
    YRT_TRY(ds.prim_attrs.alloc(sizeof(float4) * YRT_ATTR_STRIDE * (size_t)std::max(hs.n_prims, 1), device));
    YRT_TRY(ds.prim_rank.alloc(sizeof(int) * (size_t)std::max(hs.n_prims, 1), device));
    if (hs.n_prims > 0)
        k_gather_prims<<<grid_for(hs.n_prims), 256, 0, st>>>(g, blas_order.as<int>(), ds.prim_recs.as<float4>(),
                                                           ds.prim_attrs.as<float4>(), shape_has_uv.as<int>(),
                                                           ds.prim_rank_in.as<int>(), ds.prim_rank.as<int>());

    // ---- TLAS over active instances ----
    int na = ds.n_active;
    DevBuf ilo, ihi, tl_seg_of, tl_seg_first, tlas_order;
    YRT_TRY(ilo.alloc(sizeof(float4) * (size_t)std::max(na, 1), device));
    YRT_TRY(ihi.alloc(sizeof(float4) * (size_t)std::max(na, 1), device));
    YRT_TRY(tl_seg_of.alloc(sizeof(int) * (size_t)std::max(na, 1), device));
    YRT_CUDA(cudaMemsetAsync(tl_seg_of.p, 0, sizeof(int) * (size_t)std::max(na, 1), st));
    int sf[2] = {0, na};
    YRT_TRY(tl_seg_first.upload(sf, sizeof(sf), device, st));
    if (na > 0)
        k_inst_boxes<<<grid_for(na), 256, 0, st>>>(na, ds.active_inst.as<int>(), ds.inst_frame.as<float>(), ds.inst_shape.as<int>(),
                                                 ds.shape_box_lo.as<int>(), ds.shape_box_hi.as<int>(), ilo.as<float4>(), ihi.as<float4>());
    LbvhOut to;
    to.nodes = ds.nodes.as<float4>() + YRT_NODE_STRIDE * (size_t)nb_int; to.ref_offset = nb_int; to.size_bits = size_bits_tlas; to.rotate_rounds = rotate_tlas; to.rotate_pairs = rotate_pairs_tlas; to.seg_root = &ds.tlas_seg_root; to.seg_depth = &ds.tlas_seg_depth; to.seg_need = &ds.tlas_seg_need; to.order = &tlas_order;
    to.seg_box_lo = nullptr; to.seg_box_hi = nullptr;
    YRT_TRY(lbvh_build(device, st, na, 1, ilo.as<float4>(), ihi.as<float4>(), tl_seg_of.as<int>(), tl_seg_first.as<int>(), leaf_tlas, to));
    YRT_TRY(ds.inst_recs.alloc(sizeof(float4) * 4 * (size_t)std::max(na, 1), device));
    YRT_TRY(ds.inst_rank.alloc(sizeof(int) * (size_t)std::max(na, 1), device));
    if (na > 0)
        k_inst_recs<<<grid_for(na), 256, 0, st>>>(na, tlas_order.as<int>(), ds.active_inst.as<int>(), ds.inst_frame.as<float>(),
                                                ds.inst_shape.as<int>(), ds.inst_mat.as<int>(), ds.shape_kind.as<int>(),
                                                ds.blas_seg_root.as<int>(), ds.inst_recs.as<float4>(),
                                                ds.inst_rank_in.as<int>(), ds.inst_rank.as<int>());
    YRT_CUDA(cudaGetLastError());
    YRT_CUDA(cudaEventRecord(e1, st));
    YRT_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    YRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    ds.build_us = ms * 1000.f;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);

    // depths / root back to the host
    std::vector<int> bdepth(std::max(hs.n_shapes, 1), 0), bneed(std::max(hs.n_shapes, 1), 0);
    int tdepth = 0, tneed = 0, troot = YRT_REF_SENTINEL;
    YRT_CUDA(cudaMemcpy(bdepth.data(), ds.blas_seg_depth.p, sizeof(int) * bdepth.size(), cudaMemcpyDeviceToHost));
    YRT_CUDA(cudaMemcpy(bneed.data(), ds.blas_seg_need.p, sizeof(int) * bneed.size(), cudaMemcpyDeviceToHost));
    YRT_CUDA(cudaMemcpy(&tdepth, ds.tlas_seg_depth.p, sizeof(int), cudaMemcpyDeviceToHost));
    YRT_CUDA(cudaMemcpy(&tneed, ds.tlas_seg_need.p, sizeof(int), cudaMemcpyDeviceToHost));
    YRT_CUDA(cudaMemcpy(&troot, ds.tlas_seg_root.p, sizeof(int), cudaMemcpyDeviceToHost));
    ds.blas_depth = *std::max_element(bdepth.begin(), bdepth.end());
    ds.tlas_depth = tdepth;
    ds.n_blas_nodes = hs.n_prims > 1 ? hs.n_prims - 1 : 0;
    ds.n_tlas_nodes = na > 1 ? na - 1 : 0;
    if (na > 0 && troot == YRT_REF_SENTINEL) { set_error("internal: TLAS root not found"); return YRT_ERR_CUDA; }
    // entries below the TLAS root + entries below the deepest BLAS root + guard + sentinel + the "rest of a TLAS leaf" entry
    ds.stack_need = tneed + *std::max_element(bneed.begin(), bneed.end()) + 3;
    if (ds.stack_need > YRT_STACK_CAP) {
        set_error("traversal tree too deep for the stack (%d entries needed: tlas %d + blas %d levels, capacity %d)", ds.stack_need, ds.tlas_depth, ds.blas_depth, YRT_STACK_CAP);
        return YRT_ERR_UNSUPPORTED;
    }

    SceneView& v = ds.view;
    v.nodes = ds.nodes.as<float4>();
    v.inst_recs = ds.inst_recs.as<float4>();
    v.prim_recs = ds.prim_recs.as<float4>();
    v.prim_attrs = ds.prim_attrs.as<float4>();
    v.mat_recs = ds.mat_recs.as<float4>();
    v.light_recs = ds.light_recs.as<float4>();
    v.tex_rgba8 = ds.tex.as<uint8_t>();
    v.tex_info = ds.tex_info.as<int4>();
    v.srgb_lut = ds.lut.as<float>();
    v.inst_rank = ds.inst_rank.as<int>();
    v.prim_rank = ds.prim_rank.as<int>();
    v.tlas_root = troot;
    v.n_lights = (int)hs.light_inst.size();
    v.n_active_instances = na;
    return YRT_OK;
}

}  // namespace yrt
