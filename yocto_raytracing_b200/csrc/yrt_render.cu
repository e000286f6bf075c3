// yrt_render.cu — the per-frame wavefront: persistent closest-hit / any-hit traversal kernels,
// the shade kernel and the per-pixel resolve (sm_100a).
//
// Replaces raytrace() + shade() (src/raytrace.cpp:88-254) and the traversal they call
// (src/scene.cpp:386-505).  One "slot" is one camera sample ((pixel, jj, ii), raytrace.cpp:232-241);
// a batch is a run of image rows, all samples.  Per batch:
//   k_trace_closest<primary>  rays generated in-kernel (eval_camera) -> hit + world position
//   wave loop (depth = reflection recursion level):
//     k_trace_any_lights      one hit per lane, lights walked in order (a warp traces towards one light at a time)
//     k_shade                 Blinn-Phong / hair shading in light order; reflective hits push
//                             {c, kr, la} on a per-slot stack and enqueue the mirror ray (compacted
//                             queue); finished paths unwind the stack and write their radiance
//     k_trace_closest<queue>  next wave's rays; the wave's ray count stays in device memory (the host never waits for it)
//   k_resolve                 ordered per-pixel sum over (jj,ii), divide by N*N, alpha = 1; in a multi-GPU frame it stores
//                             straight into rank 0's frame (peer memory) — the gather is fused into it
// A rank's share of a frame runs as two such pipelines on two streams (kernel tails overlap).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "yrt_internal.h"
#include "yrt_shade.cuh"
#include "yrt_trace.cuh"
#include "yrt_work.cuh"

namespace yrt {

#ifndef YRT_DEFAULT_STREAMS
#define YRT_DEFAULT_STREAMS 2   /* pipelines per frame (see render_rows_device) */
#endif
#ifndef YRT_COUNTERS
#define YRT_COUNTERS 0          /* 1: the traversal kernels count their own per-ray work (node visits, box / element tests, instance entries) — a separate build of the library, never the timed one */
#endif
#define TRACE_THREADS 128
#ifndef TRACE_MIN_BLOCKS
#define TRACE_MIN_BLOCKS 8   /* resident CTAs per SM the compiler must allow (64 registers per thread) */
#endif
#ifndef TRACE_MIN_BLOCKS_ANY
#define TRACE_MIN_BLOCKS_ANY 8
#endif
#define STACK_INTS_CLOSEST (YRT_STACK_CAP * (YRT_POP_CULL ? 2 : 1))
#define STACK_INTS_ANY YRT_STACK_CAP

struct BatchParams {
    camera_k cam;
    vec3 amb;
    int width, height, samples, spp;
    int lr0;                      // first packed local row of the batch
    int tile_rows, rank, world;
    size_t cap_slots;             // stride of per-light / per-depth planes
    unsigned m_spp, m_width, m_samples;   // floor(2^32 / d) for the three divisors of slot_to_sample (udiv_by)
    GridRef cam_grid;             // apex grid of this frame's camera (nx = 0: none), read by the primary rays only
};

// n / d and n % d for a divisor known on the host: with m = floor(2^32 / d) the estimate umulhi(n, m) is the quotient or one
// below it (n / d - n m / 2^32 = n (2^32 - m d) / (d 2^32) < n / 2^32 < 1), so one compare settles it — 5 instructions
// where the generic 32-bit division takes ~20 (the ray set-up of k_trace_closest<primary> has four of them per ray)
__host__ __device__ __forceinline__ unsigned udiv_magic(unsigned d) { return d <= 1u ? 0xffffffffu : (unsigned)(0x100000000ull / d); }
__device__ __forceinline__ unsigned udiv_by(unsigned n, unsigned d, unsigned m, unsigned& rem) {
    unsigned q = d <= 1u ? n : __umulhi(n, m);
    unsigned r = n - q * d;
    if (r >= d) { q++; r -= d; }
    rem = r;
    return q;
}

// slot -> (i, j, ii, jj)
__device__ __forceinline__ void slot_to_sample(const BatchParams& bp, unsigned slot, int& i, int& j, int& ii, int& jj) {
    unsigned s, ui, uii;
    const unsigned pix = udiv_by(slot, (unsigned)bp.spp, bp.m_spp, s);
    const unsigned lr = udiv_by(pix, (unsigned)bp.width, bp.m_width, ui);
    i = (int)ui;
    j = bp.world == 1 ? bp.lr0 + (int)lr : global_row(bp.lr0 + (int)lr, bp.tile_rows, bp.rank, bp.world);   // (one rank: its packed rows are the image's rows)
    jj = (int)udiv_by(s, (unsigned)bp.samples, bp.m_samples, uii);
    ii = (int)uii;
}

// next 32 work items of the warp: false = the launch has no work left; `alive` = this lane has an item
__device__ __forceinline__ bool warp_next(const WorkDist& wd, unsigned n_items, int lane, unsigned& item, bool& alive) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(wd.counter, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n_items) return false;
    item = base + lane;
    alive = item < n_items;
    return true;
}
__device__ __forceinline__ unsigned item_count(const WorkDist& wd) { return wd.n_items_dev ? min(*wd.n_items_dev, wd.n_items) : wd.n_items; }

// ---- per-ray work counters of the traversal kernels (-DYRT_COUNTERS=1 builds only) --------------------------------
// dctr = YRT_DCTR_WORDS words per kernel class (0 primary closest, 1 queue closest, 2 any): rays, node visits, box tests,
// box tests in the instance tree, element tests, instance entries, warp-level node-loop trips x 32
#define YRT_DCTR_WORDS 8
#if YRT_COUNTERS
#define YRT_CTR_DECL TraceCounters tc_ = {0, 0, 0, 0, 0, 0, 0, 0}; unsigned long long rays_ = 0
#define YRT_CTR_PTR (&tc_)
__device__ __forceinline__ void flush_counters(unsigned long long* dctr, const TraceCounters& tc, unsigned long long rays) {
    unsigned long long v[6] = {rays, (unsigned long long)tc.node_visits, (unsigned long long)tc.box_tests, (unsigned long long)tc.tlas_box_tests,
                               (unsigned long long)tc.prim_tests, (unsigned long long)tc.inst_entries};
    for (int k = 0; k < 6; k++) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(dctr + k, x);
    }
}
#else
#define YRT_CTR_DECL
#define YRT_CTR_PTR ((TraceCounters*)nullptr)
#endif

// ---- closest hit --------------------------------------------------------------------------
// PRIMARY: slot = work index, ray from the camera. Otherwise slot = act[idx] (or idx) and the ray
// comes from ray_o/ray_d (o.xyz|tmin, d.xyz|tmax).
template <bool PRIMARY>
__global__ void __launch_bounds__(TRACE_THREADS, TRACE_MIN_BLOCKS) k_trace_closest(SceneView sv, BatchParams bp, const int* __restrict__ act,
                                                                 const float4* __restrict__ ray_o,
                                                                 const float4* __restrict__ ray_d, float4* __restrict__ hit_out,
                                                                 float4* __restrict__ P_out, WorkDist wd, unsigned long long* dctr) {
    const int lane = threadIdx.x & 31;
    const unsigned n_items = item_count(wd);
    int stack[STACK_INTS_CLOSEST];
    YRT_CTR_DECL;
    for (;;) {
        unsigned idx = 0;
        bool alive = false;
        if (!warp_next(wd, n_items, lane, idx, alive)) break;
        if (!alive) continue;
        unsigned slot;
        ray3 ray;
        int i = 0, j = 0;
        if (PRIMARY) {
            slot = idx;
            int ii, jj;
            slot_to_sample(bp, slot, i, j, ii, jj);
            float u, v;
            sample_uv(i, j, ii, jj, bp.samples, bp.width, bp.height, u, v);
            ray = eval_camera(bp.cam, u, v);
        } else {
            slot = act ? (unsigned)act[idx] : idx;
            float4 o = ray_o[slot], d = ray_d[slot];
            ray.o = xyz(o); ray.d = xyz(d); ray.tmin = o.w; ray.tmax = d.w;
        }
        HitRec h;
        if (PRIMARY) trace_camera_ray(sv, bp.cam_grid, ray, i, j, h, stack, YRT_CTR_PTR);
        else trace_ray<false>(sv, ray, h, stack, YRT_CTR_PTR);
#if YRT_COUNTERS
        rays_++;
#endif
        float4 P = mk4(0.f, 0.f, 0.f, h.dist);
        if (h.si >= 0) {
            int kind;
            vec3 p = eval_hit_pos(sv, h.si, h.prim, h.w1, h.w2, kind);
            P.x = p.x; P.y = p.y; P.z = p.z;
        }
        hit_out[slot] = mk4(int_as_float(h.si), int_as_float(h.prim), h.w1, h.w2);
        P_out[slot] = P;
    }
#if YRT_COUNTERS
    if (dctr) flush_counters(dctr, tc_, rays_);
#else
    (void)dctr;
#endif
}

// ---- any hit: shadow rays ------------------------------------------------------------------------
// One work item per hit; the lane walks the lights in order, so the 32 lanes of a warp (neighbouring
// samples) trace towards the SAME light at the same time, and the hit record / position are read once
// per hit instead of once per (hit, light).
__global__ void __launch_bounds__(TRACE_THREADS, TRACE_MIN_BLOCKS_ANY) k_trace_any_lights(SceneView sv, size_t cap_slots, const int* __restrict__ act,
                                                                    const float4* __restrict__ hit, const float4* __restrict__ P,
                                                                    unsigned* __restrict__ vis, WorkDist wd, unsigned long long* dctr, const LightGrids lg) {
    const int lane = threadIdx.x & 31;
    const unsigned n_items = item_count(wd);
    int stack[STACK_INTS_ANY];
    YRT_CTR_DECL;
    for (;;) {
        unsigned a = 0;
        bool alive = false;
        if (!warp_next(wd, n_items, lane, a, alive)) break;
        if (!alive) continue;
        unsigned slot = act ? (unsigned)act[a] : a;
        float4 h = hit[slot];
        if (float_as_int(h.x) < 0) continue;   // a miss casts no shadow rays: shade() returns before the light loop (raytrace.cpp:93)
        vec3 p = xyz(P[slot]);
        // visibility of the lights as bit masks: word w of a slot holds lights 32 w .. 32 w + 31 (one store per hit for up to 32 lights)
        unsigned vm = 0u;
        for (int k = 0; k < sv.n_lights; k++) {
            vec3 l, ke;
            float r;
            light_vector(sv, k, p, l, r, ke);
            ray3 sr = shadow_ray(p, l, r);
            HitRec hr;
            const bool occ = trace_shadow_ray(sv, lg, k, sr, hr, stack, YRT_CTR_PTR);
#if YRT_COUNTERS
            rays_++;
#endif
            if (!occ) vm |= 1u << (k & 31);
            if ((k & 31) == 31 || k == sv.n_lights - 1) {
                vis[(size_t)(k >> 5) * cap_slots + slot] = vm;
                vm = 0u;
            }
        }
    }
#if YRT_COUNTERS
    if (dctr) flush_counters(dctr, tc_, rays_);
#else
    (void)dctr;
#endif
}

// any hit on explicit rays (intersect_any, scene.cpp:489)
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_any_rays(SceneView sv, const float4* __restrict__ ray_o,
                                                                  const float4* __restrict__ ray_d, uint8_t* __restrict__ occ_out,
                                                                  WorkDist wd) {
    const int lane = threadIdx.x & 31;
    const unsigned n_items = item_count(wd);
    int stack[STACK_INTS_ANY];
    for (;;) {
        unsigned idx = 0;
        bool alive = false;
        if (!warp_next(wd, n_items, lane, idx, alive)) break;
        if (!alive) continue;
        float4 o = ray_o[idx], d = ray_d[idx];
        ray3 ray;
        ray.o = xyz(o); ray.d = xyz(d); ray.tmin = o.w; ray.tmax = d.w;
        HitRec hr;
        occ_out[idx] = trace_ray<true>(sv, ray, hr, stack, nullptr) ? 1 : 0;
    }
}

// ---- scenes with non-rigid instance frames: the same three kernels over the reference's own instance tree -----------------
// (RefTlas in yrt_scene.cuh, trace_ray_ref in yrt_trace.cuh).  Separate kernels rather than a switch in the ones above: those are
// tuned to the register (64 registers, 8 CTAs per SM) and must not change for a case the reference's own scenes never reach.
template <bool PRIMARY>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace_closest_ref(SceneView sv, RefTlas rt, BatchParams bp, const int* __restrict__ act,
                                                                      const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                                                      float4* __restrict__ hit_out, float4* __restrict__ P_out, WorkDist wd) {
    const int lane = threadIdx.x & 31;
    const unsigned n_items = item_count(wd);
    int stack[STACK_INTS_CLOSEST];
    int tstack[YRT_REF_TLAS_STACK];
    for (;;) {
        unsigned idx = 0;
        bool alive = false;
        if (!warp_next(wd, n_items, lane, idx, alive)) break;
        if (!alive) continue;
        unsigned slot;
        ray3 ray;
        if (PRIMARY) {
            slot = idx;
            int i, j, ii, jj;
            slot_to_sample(bp, slot, i, j, ii, jj);
            float u, v;
            sample_uv(i, j, ii, jj, bp.samples, bp.width, bp.height, u, v);
            ray = eval_camera(bp.cam, u, v);
        } else {
            slot = act ? (unsigned)act[idx] : idx;
            float4 o = ray_o[slot], d = ray_d[slot];
            ray.o = xyz(o); ray.d = xyz(d); ray.tmin = o.w; ray.tmax = d.w;
        }
        HitRec h;
        trace_ray_ref<false>(sv, rt, ray, h, stack, tstack, nullptr);
        float4 P = mk4(0.f, 0.f, 0.f, h.dist);
        if (h.si >= 0) {
            int kind;
            vec3 p = eval_hit_pos(sv, h.si, h.prim, h.w1, h.w2, kind);
            P.x = p.x; P.y = p.y; P.z = p.z;
        }
        hit_out[slot] = mk4(int_as_float(h.si), int_as_float(h.prim), h.w1, h.w2);
        P_out[slot] = P;
    }
}

__global__ void __launch_bounds__(TRACE_THREADS) k_trace_any_lights_ref(SceneView sv, RefTlas rt, size_t cap_slots, const int* __restrict__ act,
                                                                         const float4* __restrict__ hit, const float4* __restrict__ P,
                                                                         unsigned* __restrict__ vis, WorkDist wd) {
    const int lane = threadIdx.x & 31;
    const unsigned n_items = item_count(wd);
    int stack[STACK_INTS_ANY];
    int tstack[YRT_REF_TLAS_STACK];
    for (;;) {
        unsigned a = 0;
        bool alive = false;
        if (!warp_next(wd, n_items, lane, a, alive)) break;
        if (!alive) continue;
        unsigned slot = act ? (unsigned)act[a] : a;
        float4 h = hit[slot];
        if (float_as_int(h.x) < 0) continue;
        vec3 p = xyz(P[slot]);
        unsigned vm = 0u;
        for (int k = 0; k < sv.n_lights; k++) {
            vec3 l, ke;
            float r;
            light_vector(sv, k, p, l, r, ke);
            ray3 sr = shadow_ray(p, l, r);
            HitRec hr;
            if (!trace_ray_ref<true>(sv, rt, sr, hr, stack, tstack, nullptr)) vm |= 1u << (k & 31);
            if ((k & 31) == 31 || k == sv.n_lights - 1) {
                vis[(size_t)(k >> 5) * cap_slots + slot] = vm;
                vm = 0u;
            }
        }
    }
}

__global__ void __launch_bounds__(TRACE_THREADS) k_trace_any_rays_ref(SceneView sv, RefTlas rt, const float4* __restrict__ ray_o,
                                                                       const float4* __restrict__ ray_d, uint8_t* __restrict__ occ_out, WorkDist wd) {
    const int lane = threadIdx.x & 31;
    const unsigned n_items = item_count(wd);
    int stack[STACK_INTS_ANY];
    int tstack[YRT_REF_TLAS_STACK];
    for (;;) {
        unsigned idx = 0;
        bool alive = false;
        if (!warp_next(wd, n_items, lane, idx, alive)) break;
        if (!alive) continue;
        float4 o = ray_o[idx], d = ray_d[idx];
        ray3 ray;
        ray.o = xyz(o); ray.d = xyz(d); ray.tmin = o.w; ray.tmax = d.w;
        HitRec hr;
        occ_out[idx] = trace_ray_ref<true>(sv, rt, ray, hr, stack, tstack, nullptr) ? 1 : 0;
    }
}

// ---- shade ------------------------------------------------------------------------------------
struct FrameCounters { unsigned long long hits, reflections, misses, truncated, max_depth; };   // truncated: mirror bounces dropped at the depth cap

struct ShadeBuffers {
    const float4* hit;
    float4 *ray_o, *ray_d, *rad;
    float4* const* pstack;        // one plane of {c, kr, la} per recursion level (device table of device pointers)
    int* next_act;
    int* next_count;
    FrameCounters* fc;
};

// shade() of one sample whose shadow rays are known (vis(k) = light k unoccluded): writes the radiance of a finished
// path, or pushes {c, kr, la} and the mirror ray of a path that goes on.  Returns spawn; is_hit says whether the ray hit.
template <class VisFn>
__device__ __forceinline__ bool shade_slot(const SceneView& sv, const BatchParams& bp, const ShadeBuffers& sb, int depth, int max_depth,
                                           unsigned slot, const float4& h, const float* lut, VisFn vis, bool& is_hit, bool& truncated) {
    int si = float_as_int(h.x);
    bool spawn = false;
    vec3 value = mk3(0.f, 0.f, 0.f);          // miss: {0,0,0,1}, raytrace.cpp:93
    const size_t cap = bp.cap_slots;
    is_hit = si >= 0;
    if (is_hit) {
        vec3 ro = depth == 0 ? bp.cam.frame.o : xyz(sb.ray_o[slot]);
        vec3 c, kr, la;
        ray3 rr;
        spawn = shade_hit(sv, si, float_as_int(h.y), h.z, h.w, ro, bp.amb, lut, vis, true, value, c, kr, la, rr);
        if (spawn && depth + 1 >= max_depth) {
            // depth cap (the reference recurses without a bound, raytrace.cpp:190-204): the mirror term is dropped, the path
            // closes with c + la, and the frame statistics say how often that happened
            spawn = false;
            truncated = true;
            value = c + la;
        }
        if (spawn) {
            float4* f = sb.pstack[depth] + (size_t)slot * 3;
            f[0] = mk4(c.x, c.y, c.z, 0.f);
            f[1] = mk4(kr.x, kr.y, kr.z, 0.f);
            f[2] = mk4(la.x, la.y, la.z, 0.f);
            sb.ray_o[slot] = mk4(rr.o.x, rr.o.y, rr.o.z, rr.tmin);
            sb.ray_d[slot] = mk4(rr.d.x, rr.d.y, rr.d.z, rr.tmax);
        }
    }
    if (!spawn) {
        // unwind the recursion: each level closes with c + col*kr, then + la (raytrace.cpp:203,206)
        for (int d = depth - 1; d >= 0; d--) {
            const float4* f = sb.pstack[d] + (size_t)slot * 3;
            value = combine_reflection(xyz(f[0]), value, xyz(f[1]), xyz(f[2]));
        }
        sb.rad[slot] = mk4(value.x, value.y, value.z, 1.0f);
    }
    return spawn;
}

// Compacted queue of the next wave + frame counters, aggregated per BLOCK: the warps leave their counts in shared memory,
// one thread adds the block's totals to the global counters (one atomic per counter and block instead of one per warp:
// a million same-address atomics per frame serialise in L2) and publishes the block's base in the next wave's queue;
// queue order = slot order inside a block.  Every thread of the block must call it (it synchronises the block).
__device__ __forceinline__ void shade_epilogue(const ShadeBuffers& sb, int depth, bool valid, bool is_hit, bool spawn, bool truncated, unsigned slot) {
    __shared__ int s_cnt[8][4];      // per warp: spawned, hits, valid, truncated
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m_spawn = __ballot_sync(0xffffffffu, spawn);
    const unsigned m_trunc = __ballot_sync(0xffffffffu, truncated);
    const unsigned m_hit = __ballot_sync(0xffffffffu, is_hit);
    const unsigned m_valid = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) { s_cnt[warp][0] = __popc(m_spawn); s_cnt[warp][1] = __popc(m_hit); s_cnt[warp][2] = __popc(m_valid); s_cnt[warp][3] = __popc(m_trunc); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t[4] = {0, 0, 0, 0};
        for (int w = 0; w < 8; w++)
            for (int k = 0; k < 4; k++) t[k] += s_cnt[w][k];
        s_base = t[0] ? atomicAdd(sb.next_count, t[0]) : 0;
        if (t[1]) atomicAdd(&sb.fc->hits, (unsigned long long)t[1]);
        if (t[0]) atomicAdd(&sb.fc->reflections, (unsigned long long)t[0]);
        if (t[2] - t[1]) atomicAdd(&sb.fc->misses, (unsigned long long)(t[2] - t[1]));
        if (t[3]) atomicAdd(&sb.fc->truncated, (unsigned long long)t[3]);
        if (depth > 0 && t[2]) atomicMax(&sb.fc->max_depth, (unsigned long long)(depth + 1));
    }
    __syncthreads();
    if (spawn) {
        int off = s_base;
        for (int w = 0; w < warp; w++) off += s_cnt[w][0];
        sb.next_act[off + __popc(m_spawn & ((1u << lane) - 1u))] = (int)slot;
    }
    __syncthreads();     // s_cnt / s_base are reused by the next round of the grid-stride loop
}

// one thread per active sample; n_act_dev (if not null) holds the number of active samples of this wave
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, SHADE_MIN_BLOCKS) k_shade(SceneView sv, BatchParams bp, ShadeBuffers sb, int depth, int max_depth,
                                                  const int* __restrict__ act, unsigned n_act, const unsigned* __restrict__ n_act_dev,
                                                  const unsigned* __restrict__ vis) {
    if (n_act_dev) n_act = min(n_act, *n_act_dev);
    if (blockIdx.x * blockDim.x >= n_act) return;
    __shared__ float lut[256];
    lut[threadIdx.x] = sv.srgb_lut[threadIdx.x];
    __syncthreads();
    // (grid-stride: waves whose size is only known on the device are launched with a bounded grid)
    for (unsigned base = blockIdx.x * blockDim.x; base < n_act; base += gridDim.x * blockDim.x) {
        unsigned a = base + threadIdx.x;
        bool valid = a < n_act;
        bool is_hit = false, spawn = false, truncated = false;
        unsigned slot = 0;
        if (valid) {
            slot = act ? (unsigned)act[a] : a;
            float4 h = sb.hit[slot];
            const unsigned* vrow = vis + slot;
            const size_t cap = bp.cap_slots;
            const unsigned v0 = vrow[0];      // lights 0..31: requested together with the hit record, used after eval_hit
            spawn = shade_slot(sv, bp, sb, depth, max_depth, slot, h, lut,
                               [&](int k) { return (((k < 32 ? v0 : vrow[(size_t)(k >> 5) * cap]) >> (k & 31)) & 1u) != 0u; }, is_hit, truncated);
        }
        shade_epilogue(sb, depth, valid, is_hit, spawn, truncated, slot);
    }
}

// scatter != 0: `out` is the FULL frame (possibly another GPU's memory mapped over NVLink) and every pixel goes to
// its final position — the gather of the multi-GPU path fused into the resolve (no packed rows, no copy, no unpack)
__global__ void __launch_bounds__(256) k_resolve(const float4* __restrict__ rad, float4* __restrict__ out, int n_pix, int spp,
                                                 size_t out_pix0, int scatter, int width, int lr0, int tile_rows, int rank, int world) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    const float4* r = rad + (size_t)p * spp;
    float x = 0.f, y = 0.f, z = 0.f;
    for (int s = 0; s < spp; s++) {   // (jj outer, ii inner) = slot order
        float4 v = r[s];
        x += v.x; y += v.y; z += v.z;
    }
    float d = (float)spp;             // float(samples * samples)
    size_t o = out_pix0 + p;
    if (scatter) {
        int lr = p / width, i = p - lr * width;
        o = (size_t)global_row(lr0 + lr, tile_rows, rank, world) * width + i;
    }
    out[o] = mk4(x / d, y / d, z / d, 1.0f);
}

// ---- parity hook: hit records -> (instance, shape, element) ids ----------------------------------
__global__ void k_hit_ids(SceneView sv, const float4* __restrict__ hit, const float4* __restrict__ P, int n, int* __restrict__ ids,
                          float* __restrict__ dist, float* __restrict__ uv) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float4 h = hit[s];
    int si = float_as_int(h.x);
    int inst = -1, shape = -1, ei = -1;
    float d = 0.f, a = 0.f, b = 0.f;
    if (si >= 0) {
        const float4* ir = sv.inst_recs + 4 * (size_t)si;
        inst = float_as_int(ir[1].w);
        shape = float_as_int(ir[3].w) & 0x0fffffff;
        ei = float_as_int(sv.prim_recs[3 * (size_t)float_as_int(h.y)].w);
        d = P[s].w; a = h.z; b = h.w;
    }
    ids[3 * (size_t)s] = inst; ids[3 * (size_t)s + 1] = shape; ids[3 * (size_t)s + 2] = ei;
    if (dist) dist[s] = d;
    if (uv) { uv[2 * (size_t)s] = a; uv[2 * (size_t)s + 1] = b; }
}

__global__ void k_unpack_rows(const float4* __restrict__ packed, float4* __restrict__ full, int width, int n_rows, int tile_rows,
                              int rank, int world) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)n_rows * width;
    if (t >= n) return;
    int lr = (int)(t / width), i = (int)(t - (size_t)lr * width);
    int j = global_row(lr, tile_rows, rank, world);
    full[(size_t)j * width + i] = packed[t];
}

// tonemap(hdr, 0, false) of src/image.cpp:55-78: powf(2,0)=1 scale, powf(x,1/2.2f), clamp, truncate
__global__ void k_tonemap(const float4* __restrict__ in, uchar4* __restrict__ out, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float4 h = in[t];
    const float e = 1.0f;   // powf(2, exposure) with exposure 0
    float g = 1 / 2.2f;
    float x = powf(h.x * e, g), y = powf(h.y * e, g), z = powf(h.z * e, g), w = h.w;
    uchar4 o;
    o.x = (unsigned char)(rclamp(x, 0.0f, 1.0f) * 255);
    o.y = (unsigned char)(rclamp(y, 0.0f, 1.0f) * 255);
    o.z = (unsigned char)(rclamp(z, 0.0f, 1.0f) * 255);
    o.w = (unsigned char)(rclamp(w, 0.0f, 1.0f) * 255);
    out[t] = o;
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
camera_k make_camera_k(const yrt_camera* cam) {
    camera_k c;
    const float* f = cam->frame;
    c.frame.x = mk3(f[0], f[1], f[2]); c.frame.y = mk3(f[3], f[4], f[5]); c.frame.z = mk3(f[6], f[7], f[8]); c.frame.o = mk3(f[9], f[10], f[11]);
    // raytrace.cpp:21-22 on the host: tan() of a float argument resolves to tanf
    c.h = 2.0f * cam->focus * tanf(cam->fovy / 2.0f);
    c.w = c.h * cam->aspect;
    c.focus = cam->focus;
    return c;
}

static int env_int(const char* name, int def) {
    const char* e = getenv(name);
    return e ? atoi(e) : def;
}

struct PhaseTimer {
    struct Span { int cat; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    bool on = false;
    bool deferred = false;        // between yrt_stats_begin and yrt_stats_end: record, never synchronise
    int frames = 0;               // frames recorded since the last collect
    int64_t primary = 0;          // primary rays of those frames
    cudaStream_t st = nullptr;      // stream the next span is recorded on
    cudaStream_t main = nullptr;    // the frame's stream (what collect waits for)
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void begin(int cat) { if (!on) return; Span s; s.cat = cat; s.a = get(); s.b = get(); cudaEventRecord(s.a, st); spans.push_back(s); }
    void end() { if (!on) return; cudaEventRecord(spans.back().b, st); }
    int64_t extra_launches = 0;   // kernels launched without an event span (the camera grid build on its own stream)
    void collect(float out[5], int cnt[5], int64_t& launches) {
        for (int i = 0; i < 5; i++) { out[i] = 0.f; cnt[i] = 0; }
        launches = extra_launches;
        extra_launches = 0;
        for (auto& s : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s.a, s.b);
            out[s.cat] += ms; cnt[s.cat]++;
            if (s.cat != 4) launches++;     // CAT_FRAME spans bracket a whole frame, they are not launches
            pool.push_back(s.a); pool.push_back(s.b);
        }
        spans.clear();
    }
    ~PhaseTimer() { for (auto e : pool) cudaEventDestroy(e); for (auto& s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); } }
};
enum { CAT_CLOSEST = 0, CAT_ANY = 1, CAT_SHADE = 2, CAT_OTHER = 3, CAT_FRAME = 4 };

static int batch_rows_for(const RenderParams& rp, int n_lights, int own_rows, bool reflective);
#define YRT_MAX_WAVES 64                                   /* the reflection depth cap: default and upper limit of YRT_MAX_DEPTH (the reference recurses without a bound; a mirror chain of 64 bounces carries kr^64 of its light) */
#define STATS_WAVE_COUNTS_OFFSET 64                          /* bytes: FrameCounters, then one active-sample count per wave */
#define STATS_BYTES (STATS_WAVE_COUNTS_OFFSET + sizeof(unsigned) * (YRT_MAX_WAVES + 2))

static int ensure_workspace(DevScene& ds, Workspace& w, size_t slots, int n_lights, int depth_cap, bool reflective) {
    int dev = ds.device;
    if (slots > w.cap_slots || n_lights > w.cap_lights || (reflective ? depth_cap : 0) > w.cap_depth) {
        size_t cs = std::max(slots, w.cap_slots);
        int cl = std::max(n_lights, w.cap_lights);
        int cd = std::max(reflective ? depth_cap : 0, w.cap_depth);
        YRT_TRY(w.hit.alloc(sizeof(float4) * cs, dev));
        YRT_TRY(w.P.alloc(sizeof(float4) * cs, dev));
        YRT_TRY(w.rad.alloc(sizeof(float4) * cs, dev));
        YRT_TRY(w.vis.alloc(sizeof(unsigned) * (size_t)((std::max(cl, 1) + 31) / 32) * cs, dev));
        if (cd > 0) {
            YRT_TRY(w.ray_o.alloc(sizeof(float4) * cs, dev));
            YRT_TRY(w.ray_d.alloc(sizeof(float4) * cs, dev));
            for (DevBuf& b : w.pstack_lvl) b.release();   // per-level planes are (re)allocated on demand at the new size
            w.pstack_levels = 0;
            YRT_TRY(w.act0.alloc(sizeof(int) * cs, dev));
            YRT_TRY(w.act1.alloc(sizeof(int) * cs, dev));
        }
        w.cap_slots = cs; w.cap_lights = cl; w.cap_depth = cd;
    }
    YRT_TRY(w.counters.alloc(sizeof(unsigned) * YRT_WORK_BLOCKS, dev));
    YRT_TRY(w.stats.alloc(STATS_BYTES, dev));
    if (reflective && !w.h_counts) {
        YRT_CUDA(cudaHostAlloc((void**)&w.h_counts, sizeof(unsigned) * (YRT_MAX_WAVES + 2), cudaHostAllocDefault));
        for (int k = 0; k < YRT_MAX_WAVES + 2; k++) YRT_CUDA(cudaEventCreateWithFlags(&w.ev_counts[k], cudaEventDisableTiming));
    }
#if YRT_COUNTERS
    if (!ds.dctr.p) {
        YRT_TRY(ds.dctr.alloc(sizeof(unsigned long long) * 3 * YRT_DCTR_WORDS, dev));
        YRT_CUDA(cudaMemset(ds.dctr.p, 0, sizeof(unsigned long long) * 3 * YRT_DCTR_WORDS));
    }
#endif
    return YRT_OK;
}

// The {c, kr, la} plane of recursion level `level` (48 B per slot) is allocated when a wave first gets that deep — the
// host is at most one wave ahead of the counts (run_batch), so the memory follows the depth the scene really reaches
// (+ 2 levels), not the depth cap.  The table of plane pointers lives in device memory and is updated in stream order.
static int ensure_pstack_level(DevScene& ds, Workspace& w, int level, cudaStream_t st) {
    if (level < w.pstack_levels) return YRT_OK;
    YRT_TRY(w.pstack_tab.alloc(sizeof(float4*) * (YRT_MAX_WAVES + 2), ds.device));
    if (!w.h_pstack_tab) YRT_CUDA(cudaHostAlloc((void**)&w.h_pstack_tab, sizeof(float4*) * (YRT_MAX_WAVES + 2), cudaHostAllocDefault));
    for (int l = w.pstack_levels; l <= level; l++) {
        YRT_TRY(w.pstack_lvl[l].alloc(sizeof(float4) * 3 * w.cap_slots, ds.device));
        w.h_pstack_tab[l] = w.pstack_lvl[l].as<float4>();
    }
    w.pstack_levels = level + 1;
    YRT_CUDA(cudaMemcpyAsync(w.pstack_tab.p, w.h_pstack_tab, sizeof(float4*) * (size_t)w.pstack_levels, cudaMemcpyHostToDevice, st));
    return YRT_OK;
}

static int persistent_grid(DevScene& ds, const void* kernel);

// Everything the first frame of that size would otherwise pay for inside the render call: the workspace, the camera grid's
// buffers, and the frame kernels' modules (CUDA loads a kernel lazily at its first launch or attribute query).
int presize_workspace_device(DevScene& ds, int width, int height, int samples) {
    RenderParams rp;
    rp.width = width; rp.height = height; rp.samples = samples; rp.tile_rows = std::max(height, 1); rp.rank = 0; rp.world = 1;
    YRT_CUDA(cudaSetDevice(ds.device));
    int depth_cap = std::min(YRT_MAX_WAVES, std::max(1, env_int("YRT_MAX_DEPTH", YRT_MAX_WAVES)));
    size_t cap_slots = (size_t)batch_rows_for(rp, ds.view.n_lights, height, ds.has_reflective) * width * samples * samples;
    YRT_TRY(ensure_workspace(ds, ds.ws, cap_slots, ds.view.n_lights, depth_cap, ds.has_reflective));
    const bool ref = ds.ref.n_nodes > 0;
    if (!ds.grid_closest_primary) ds.grid_closest_primary = persistent_grid(ds, ref ? (const void*)k_trace_closest_ref<true> : (const void*)k_trace_closest<true>);
    if (!ds.grid_closest_queue) ds.grid_closest_queue = persistent_grid(ds, ref ? (const void*)k_trace_closest_ref<false> : (const void*)k_trace_closest<false>);
    if (!ds.grid_any) ds.grid_any = persistent_grid(ds, ref ? (const void*)k_trace_any_lights_ref : (const void*)k_trace_any_lights);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, (const void*)k_shade);
    cudaFuncGetAttributes(&fa, (const void*)k_resolve);
    cudaFuncGetAttributes(&fa, (const void*)k_tonemap);
    {   // a camera grid for a stand-in camera of this frame size: allocates the grid's buffers and loads its kernels
        camera_k c;
        c.frame.x = mk3(1.f, 0.f, 0.f); c.frame.y = mk3(0.f, 1.f, 0.f); c.frame.z = mk3(0.f, 0.f, 1.f); c.frame.o = mk3(0.f, 0.f, 0.f);
        c.h = 1.0f; c.w = (float)width / (float)std::max(height, 1); c.focus = 1.0f;
        GridRef g;
        YRT_TRY(camera_grid_enqueue(ds, c, width, height, ds.stream, &g));
        YRT_CUDA(cudaStreamSynchronize(ds.stream));
    }
    cudaGetLastError();
    return YRT_OK;
}

static int persistent_grid(DevScene& ds, const void* kernel) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRACE_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    int cap = env_int("YRT_BLOCKS_PER_SM", 0);
    if (cap > 0 && cap < per_sm) per_sm = cap;
    return ds.sm_count * per_sm;
}

// per-launch work counters, zeroed in one go and handed out in turn
struct CounterRing {
    unsigned* base; int next; int cap; cudaStream_t st;
    int init(Workspace& w, cudaStream_t s) {
        base = w.counters.as<unsigned>(); next = 0; cap = YRT_WORK_BLOCKS; st = s;
        YRT_CUDA(cudaMemsetAsync(base, 0, sizeof(unsigned) * cap, st));
        return YRT_OK;
    }
    int get(unsigned** out) {
        if (next == cap) { YRT_CUDA(cudaMemsetAsync(base, 0, sizeof(unsigned) * cap, st)); next = 0; }
        *out = base + next++;
        return YRT_OK;
    }
};

// batch of rows [lr0, lr0+nrows) of the rank's packed rows; primary hits only when primary_only
static int run_batch(DevScene& ds, Workspace& w, const RenderParams& rp, const GridRef& cam_grid, int lr0, int nrows, size_t cap_slots, float4* d_out, cudaStream_t st,
                     PhaseTimer& pt, CounterRing& ring, int depth_cap, bool reflective, bool primary_only, int pipe = -1) {
    BatchParams bp;
    bp.cam_grid = cam_grid;
    bp.cam = rp.cam; bp.amb = rp.amb; bp.width = rp.width; bp.height = rp.height; bp.samples = rp.samples;
    bp.m_spp = udiv_magic((unsigned)(rp.samples * rp.samples)); bp.m_width = udiv_magic((unsigned)rp.width); bp.m_samples = udiv_magic((unsigned)rp.samples);
    bp.spp = rp.samples * rp.samples; bp.lr0 = lr0; bp.tile_rows = rp.tile_rows; bp.rank = rp.rank; bp.world = rp.world;
    bp.cap_slots = cap_slots;
    unsigned n = (unsigned)((size_t)nrows * rp.width * bp.spp);
    int nl = ds.view.n_lights;
    const bool ref = ds.ref.n_nodes > 0;
    if (!ds.grid_closest_primary) ds.grid_closest_primary = persistent_grid(ds, ref ? (const void*)k_trace_closest_ref<true> : (const void*)k_trace_closest<true>);
    if (!ds.grid_closest_queue) ds.grid_closest_queue = persistent_grid(ds, ref ? (const void*)k_trace_closest_ref<false> : (const void*)k_trace_closest<false>);
    if (!ds.grid_any) ds.grid_any = persistent_grid(ds, ref ? (const void*)k_trace_any_lights_ref : (const void*)k_trace_any_lights);
    auto grid_of = [](int g, unsigned items) { unsigned need = (items + TRACE_THREADS - 1) / TRACE_THREADS; return (int)std::max(1u, std::min((unsigned)g, need)); };
    unsigned long long* dctr = YRT_COUNTERS ? ds.dctr.as<unsigned long long>() : nullptr;

    unsigned* ctr = nullptr;
    YRT_TRY(ring.get(&ctr));
    pt.begin(CAT_CLOSEST);
    if (ref) k_trace_closest_ref<true><<<grid_of(ds.grid_closest_primary, n), TRACE_THREADS, 0, st>>>(ds.view, ds.ref, bp, nullptr, nullptr, nullptr, w.hit.as<float4>(), w.P.as<float4>(), workdist_linear(ctr, n));
    else k_trace_closest<true><<<grid_of(ds.grid_closest_primary, n), TRACE_THREADS, 0, st>>>(ds.view, bp, nullptr, nullptr, nullptr, w.hit.as<float4>(), w.P.as<float4>(),
                                                                                          workdist_linear(ctr, n), dctr);
    pt.end();
    if (pipe >= 0 && cam_grid.nx > 0 && ds.ev_primary_done[pipe]) {     // the last reader of this frame's camera grid on this pipeline so far
        YRT_CUDA(cudaEventRecord(ds.ev_primary_done[pipe], st));
        ds.primary_recorded[pipe] = true;
    }
    if (primary_only) { YRT_CUDA(cudaGetLastError()); return YRT_OK; }

    // Wave loop.  Wave d traces the shadow rays of its active samples, shades them, and appends the mirror rays it spawns to
    // the queue of wave d + 1, whose length is counts[d + 1] in device memory: every kernel of a later wave reads its item
    // count from there, so a wave is enqueued before the previous one has run.  To know when to stop, the host reads the
    // counts back (asynchronous copy into pinned memory + event) ONE WAVE LATE: before it enqueues wave d it waits for the
    // count of wave d - 1, while the kernels of wave d - 1 are still queued behind it on the device — the GPU never idles,
    // and at most one empty wave is launched beyond the last real one (none at all for scenes without mirrors).
    FrameCounters* fc = ds.ws.stats.as<FrameCounters>();   // one set of frame counters for all pipelines
    unsigned* counts = (unsigned*)((char*)w.stats.p + STATS_WAVE_COUNTS_OFFSET);
    if (reflective) YRT_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned) * (YRT_MAX_WAVES + 2), st));
    int* act_bufs[2] = {w.act0.as<int>(), w.act1.as<int>()};
    for (int depth = 0;; depth++) {
        const int* act = depth == 0 ? nullptr : act_bufs[(depth - 1) & 1];
        const unsigned* n_dev = depth == 0 ? nullptr : counts + depth;
        if (depth > 0) {
            YRT_TRY(ring.get(&ctr));
            pt.begin(CAT_CLOSEST);
            if (ref) k_trace_closest_ref<false><<<grid_of(ds.grid_closest_queue, n), TRACE_THREADS, 0, st>>>(ds.view, ds.ref, bp, act, w.ray_o.as<float4>(), w.ray_d.as<float4>(), w.hit.as<float4>(),
                                                                                                         w.P.as<float4>(), workdist_linear(ctr, n, n_dev));
            else k_trace_closest<false><<<grid_of(ds.grid_closest_queue, n), TRACE_THREADS, 0, st>>>(ds.view, bp, act, w.ray_o.as<float4>(), w.ray_d.as<float4>(), w.hit.as<float4>(),
                                                                                               w.P.as<float4>(), workdist_linear(ctr, n, n_dev), dctr ? dctr + YRT_DCTR_WORDS : nullptr);
            pt.end();
        }
        if (reflective && depth + 1 < depth_cap) YRT_TRY(ensure_pstack_level(ds, w, depth, st));
        ShadeBuffers sb;
        sb.hit = w.hit.as<float4>(); sb.ray_o = w.ray_o.as<float4>(); sb.ray_d = w.ray_d.as<float4>(); sb.pstack = w.pstack_tab.as<float4*>();
        sb.rad = w.rad.as<float4>(); sb.next_act = reflective ? act_bufs[depth & 1] : nullptr; sb.next_count = (int*)(counts + depth + 1); sb.fc = fc;
        if (nl > 0) {
            YRT_TRY(ring.get(&ctr));
            pt.begin(CAT_ANY);
            if (ref) k_trace_any_lights_ref<<<grid_of(ds.grid_any, n), TRACE_THREADS, 0, st>>>(ds.view, ds.ref, cap_slots, act, w.hit.as<float4>(), w.P.as<float4>(), w.vis.as<unsigned>(),
                                                                                          workdist_linear(ctr, n, n_dev));
            else k_trace_any_lights<<<grid_of(ds.grid_any, n), TRACE_THREADS, 0, st>>>(ds.view, cap_slots, act, w.hit.as<float4>(), w.P.as<float4>(), w.vis.as<unsigned>(),
                                                                                 workdist_linear(ctr, n, n_dev), dctr ? dctr + 2 * YRT_DCTR_WORDS : nullptr, ds.light_grids);
            pt.end();
        }
        pt.begin(CAT_SHADE);
        k_shade<<<depth == 0 ? (n + 255) / 256 : std::min((n + 255) / 256, (unsigned)ds.sm_count * 16u), 256, 0, st>>>(ds.view, bp, sb, depth, depth_cap, act, n, n_dev, w.vis.as<unsigned>());
        pt.end();
        if (!reflective || depth + 1 >= depth_cap) break;
        YRT_CUDA(cudaMemcpyAsync(w.h_counts + depth + 1, counts + depth + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        YRT_CUDA(cudaEventRecord(w.ev_counts[depth + 1], st));
        if (depth >= 1) {   // about to enqueue wave depth + 1: the count of wave `depth` must have arrived
            YRT_CUDA(cudaEventSynchronize(w.ev_counts[depth]));
            if (w.h_counts[depth] == 0) break;
        }
    }
    int n_pix = nrows * rp.width;
    pt.begin(CAT_OTHER);
    k_resolve<<<(n_pix + 255) / 256, 256, 0, st>>>(w.rad.as<float4>(), d_out, n_pix, bp.spp, (size_t)lr0 * rp.width, rp.scatter ? 1 : 0, rp.width, lr0,
                                                  rp.tile_rows, rp.rank, rp.world);
    pt.end();
    YRT_CUDA(cudaGetLastError());
    return YRT_OK;
}

// Packed local rows [lr0, lr0 + nrows) of one rank -> their places in a row-major HOST frame (elem = bytes per pixel), as
// 2-D copies on `st`: the rows of tile k of rank r are one contiguous block of the frame ((k * world + r) * tile_rows rows
// down) and the rank's packed buffer holds its blocks back to back, so a pitched copy moves a run of full tiles at once;
// ragged ends (a batch that starts or stops inside a tile, the short last tile of the image) go row block by row block.
int copy_rows_to_host(const void* d_packed, void* h_frame, int width, int height, int tile_rows, int rank, int world, size_t elem, int lr0, int nrows,
                      cudaStream_t st) {
    const size_t row = elem * (size_t)width;
    int lr = lr0;
    const int end = lr0 + nrows;
    while (lr < end) {
        const int k = lr / tile_rows, in_tile = lr - k * tile_rows;            // k-th tile of this rank
        const int g0 = (k * world + rank) * tile_rows;                          // its first global row
        const int tile_len = std::min(tile_rows, height - g0);                  // (the image's last tile may be short)
        if (in_tile == 0 && tile_len == tile_rows && end - lr >= tile_rows) {
            // run of full tiles: stop before a tile that is cut by the batch end or by the image end
            int full = (end - lr) / tile_rows;
            while (full > 0 && ((k + full - 1) * world + rank + 1) * tile_rows > height) full--;
            if (full > 0) {
                const size_t block = row * (size_t)tile_rows;
                YRT_CUDA(cudaMemcpy2DAsync((char*)h_frame + row * (size_t)g0, block * (size_t)world, (const char*)d_packed + row * (size_t)lr, block, block, (size_t)full,
                                           cudaMemcpyDeviceToHost, st));
                lr += full * tile_rows;
                continue;
            }
        }
        const int n = std::min(tile_len - in_tile, end - lr);
        YRT_CUDA(cudaMemcpyAsync((char*)h_frame + row * (size_t)(g0 + in_tile), (const char*)d_packed + row * (size_t)lr, row * (size_t)n, cudaMemcpyDeviceToHost, st));
        lr += n;
    }
    return YRT_OK;
}

// rows [lr0, lr0 + nrows) of the rank's packed rows -> the host destinations of rp (float frame and / or tonemapped RGBA8)
static int rows_to_host(const RenderParams& rp, const float4* d_out, int lr0, int nrows, cudaStream_t st) {
    if (nrows <= 0) return YRT_OK;
    if (rp.h_rgba) YRT_TRY(copy_rows_to_host(d_out, rp.h_rgba, rp.width, rp.height, rp.tile_rows, rp.rank, rp.world, sizeof(float4), lr0, nrows, st));
    if (rp.h_ldr) {
        YRT_TRY(tonemap_launch(d_out + (size_t)lr0 * rp.width, rp.d_ldr_rows + 4 * (size_t)lr0 * rp.width, (size_t)nrows * rp.width, st));
        YRT_TRY(copy_rows_to_host(rp.d_ldr_rows, rp.h_ldr, rp.width, rp.height, rp.tile_rows, rp.rank, rp.world, 4, lr0, nrows, st));
    }
    return YRT_OK;
}

static int check_params(const RenderParams& rp) {
    if (rp.width <= 0 || rp.height <= 0 || rp.samples <= 0) { set_error("width, height and samples must be positive"); return YRT_ERR_INVALID; }
    if (rp.tile_rows <= 0 || rp.world <= 0 || rp.rank < 0 || rp.rank >= rp.world) { set_error("bad row-tile partition (tile_rows %d rank %d world %d)", rp.tile_rows, rp.rank, rp.world); return YRT_ERR_INVALID; }
    if ((long long)rp.samples * rp.samples > (1 << 20)) { set_error("samples too large"); return YRT_ERR_INVALID; }
    return YRT_OK;
}

static int batch_rows_for(const RenderParams& rp, int n_lights, int own_rows, bool reflective) {
    // slots (camera samples) per batch: a whole 1080p/16spp frame (33.2 M slots, 51 B each) when nothing reflects —
    // fewer launches and kernel tails; 8.5 M (a 1280x720, 9 spp frame) when the per-slot recursion stack
    // (48 B x depth cap per slot) must be allocated too
    long long target = env_int("YRT_BATCH_SLOTS", reflective ? 8500000 : (48 << 20));
    long long per_row = (long long)rp.width * rp.samples * rp.samples;
    if (target > 0x7fffff00ll) target = 0x7fffff00ll;
    (void)n_lights;
    long long rows = std::max(1ll, target / per_row);
    return (int)std::min<long long>(rows, std::max(own_rows, 1));
}

int render_rows_device(DevScene& ds, const RenderParams& rp, float4* d_out, cudaStream_t st, yrt_stats* stats, bool sync_for_stats) {
    YRT_TRY(check_params(rp));
    YRT_CUDA(cudaSetDevice(ds.device));
    int own = rows_owned(rp.height, rp.tile_rows, rp.rank, rp.world);
    int spp = rp.samples * rp.samples;
    int nl = ds.view.n_lights;
    bool reflective = ds.has_reflective;
    int depth_cap = std::min(YRT_MAX_WAVES, std::max(1, env_int("YRT_MAX_DEPTH", YRT_MAX_WAVES)));
    int batch_rows = batch_rows_for(rp, nl, own, reflective);
    // two pipelines: the rows are cut into (at least) two batches that run on two streams, so the ramp-up of one
    // batch's kernel fills the tail of the other's (each persistent kernel ends with ~one 32-ray task of idle SMs), and a
    // batch's rows go to the host under the other batch's kernels.  Measured on the final kernels (tools/diag_streams.py):
    // 1/8 frame 1.475 -> 1.361 ms (3 pipelines 1.372, 4: 1.375), 1/4 frame 2.706 -> 2.566, whole frame 9.964 -> 9.865 ms.
    // One pipeline only when per-kernel statistics of THIS call are asked for (event spans are unambiguous on one stream).
    const bool to_host = rp.h_rgba != nullptr || rp.h_ldr != nullptr;
    const int n_pipes = std::max(1, std::min(std::min(4, own), env_int("YRT_STREAMS", stats ? 1 : YRT_DEFAULT_STREAMS)));
    if (n_pipes > 1) batch_rows = std::min(batch_rows, (own + n_pipes - 1) / n_pipes);
    // rows that go on to the host: two batches per pipeline, so that the copy left exposed at the end of the frame (the last
    // batch's) is a quarter of the frame, not half (tools/e2e_streams.py: 10.44 -> 10.37 ms per 1080p frame end to end)
    // (only for large shares: the launches of a small share are short as it is)
    if (n_pipes > 1 && to_host && !rp.scatter && (long long)own * rp.width * spp >= (16ll << 20)) batch_rows = std::min(batch_rows, (own + 2 * n_pipes - 1) / (2 * n_pipes));
    size_t cap_slots = (size_t)batch_rows * rp.width * spp;
    YRT_TRY(ensure_workspace(ds, ds.ws, cap_slots, nl, depth_cap, reflective));
    for (int k = 1; k < n_pipes; k++) {
        YRT_TRY(ensure_workspace(ds, ds.ws_aux[k - 1], cap_slots, nl, depth_cap, reflective));
        if (!ds.aux_stream[k - 1]) YRT_CUDA(cudaStreamCreateWithFlags(&ds.aux_stream[k - 1], cudaStreamNonBlocking));
        if (!ds.ev_join[k - 1]) YRT_CUDA(cudaEventCreateWithFlags(&ds.ev_join[k - 1], cudaEventDisableTiming));
    }
    if (n_pipes > 1 && !ds.ev_fork) YRT_CUDA(cudaEventCreateWithFlags(&ds.ev_fork, cudaEventDisableTiming));

    if (!ds.timer) ds.timer = new PhaseTimer();
    PhaseTimer& pt = *ds.timer;
    pt.on = stats != nullptr || pt.deferred;
    pt.st = st;
    pt.main = st;
    if (!pt.deferred) {   // a stand-alone frame starts from clean counters; deferred frames accumulate
        for (auto& sp : pt.spans) { pt.pool.push_back(sp.a); pt.pool.push_back(sp.b); }
        pt.spans.clear();
        pt.frames = 0; pt.primary = 0; pt.extra_launches = 0;
        YRT_CUDA(cudaMemsetAsync(ds.ws.stats.p, 0, sizeof(FrameCounters), st));
    }
    pt.begin(CAT_FRAME);
    const size_t frame_span = pt.spans.size();
    CounterRing ring[4];
    YRT_TRY(ring[0].init(ds.ws, st));
    // This frame's camera grid (yrt_pgrid.cuh), rebuilt for every frame, on its own stream: it waits for the primary-ray
    // kernels of the previous frame (the only readers of the grid's buffers), so with frames enqueued back to back it runs
    // under the previous frame's shadow / shade kernels; the pipelines wait for it before their first kernel.
    GridRef cam_grid;
    if (!ds.grid_stream) {
        YRT_CUDA(cudaStreamCreateWithFlags(&ds.grid_stream, cudaStreamNonBlocking));
        YRT_CUDA(cudaEventCreateWithFlags(&ds.ev_grid, cudaEventDisableTiming));
        for (int k = 0; k < 4; k++) YRT_CUDA(cudaEventCreateWithFlags(&ds.ev_primary_done[k], cudaEventDisableTiming));
    }
    const bool own_stream = env_int("YRT_GRID_STREAM", 1) != 0;
    cudaStream_t gst = own_stream ? ds.grid_stream : st;
    for (int k = 0; k < 4; k++)
        if (ds.primary_recorded[k]) { if (own_stream) YRT_CUDA(cudaStreamWaitEvent(ds.grid_stream, ds.ev_primary_done[k], 0)); ds.primary_recorded[k] = false; }
    // (no event span around it: on its own stream the three small kernels wait for SMs the previous frame's persistent kernels
    //  hold, so a span would measure that wait; their own time is in the ncu launch lists, 39 us)
    YRT_TRY(camera_grid_enqueue(ds, rp.cam, rp.width, rp.height, gst, &cam_grid));
    if (pt.on && cam_grid.nx > 0) pt.extra_launches += 3;    // scatter, scatter (large instances), emit
    if (cam_grid.nx > 0 && own_stream) {
        YRT_CUDA(cudaEventRecord(ds.ev_grid, ds.grid_stream));
        YRT_CUDA(cudaStreamWaitEvent(st, ds.ev_grid, 0));
    }
    if (n_pipes > 1) YRT_CUDA(cudaEventRecord(ds.ev_fork, st));
    for (int k = 1; k < n_pipes; k++) {
        YRT_CUDA(cudaStreamWaitEvent(ds.aux_stream[k - 1], ds.ev_fork, 0));
        YRT_TRY(ring[k].init(ds.ws_aux[k - 1], ds.aux_stream[k - 1]));
    }
    int b = 0;
    for (int lr0 = 0; lr0 < own; lr0 += batch_rows, b++) {
        int nrows = std::min(batch_rows, own - lr0);
        int pipe = b % n_pipes;
        Workspace& w = pipe ? ds.ws_aux[pipe - 1] : ds.ws;
        pt.st = pipe ? ds.aux_stream[pipe - 1] : st;
        YRT_TRY(run_batch(ds, w, rp, cam_grid, lr0, nrows, w.cap_slots, d_out, pt.st, pt, ring[pipe], depth_cap, reflective, false, pipe));
        if (to_host && !rp.scatter && !pt.on) YRT_TRY(rows_to_host(rp, d_out, lr0, nrows, pt.st));
    }
    pt.st = st;
    for (int k = 1; k < n_pipes; k++) {
        YRT_CUDA(cudaEventRecord(ds.ev_join[k - 1], ds.aux_stream[k - 1]));
        YRT_CUDA(cudaStreamWaitEvent(st, ds.ev_join[k - 1], 0));
    }
    if (pt.on) {
        cudaEventRecord(pt.spans[frame_span - 1].b, st);
        pt.frames++;
        pt.primary += (int64_t)own * rp.width * spp;
        // with statistics the frame span is the rendering alone: the rows go to the host behind it, in one piece
        if (to_host && !rp.scatter) YRT_TRY(rows_to_host(rp, d_out, 0, own, st));
    }
    if (stats && sync_for_stats && !pt.deferred) YRT_TRY(collect_stats_device(ds, rp, stats));
    return YRT_OK;
}

int collect_stats_device(DevScene& ds, const RenderParams& rp, yrt_stats* stats) {
    (void)rp;
    if (!ds.timer) { set_error("collect_stats_device: no frame was rendered with stats"); return YRT_ERR_INVALID; }
    PhaseTimer& pt = *ds.timer;
    YRT_CUDA(cudaSetDevice(ds.device));
    YRT_CUDA(cudaStreamSynchronize(pt.main));
    FrameCounters fc;
    YRT_CUDA(cudaMemcpyAsync(&fc, ds.ws.stats.p, sizeof(fc), cudaMemcpyDeviceToHost, pt.main));
    YRT_CUDA(cudaStreamSynchronize(pt.main));
    float cat[5];
    int cnt[5];
    int64_t launches = 0;
    pt.collect(cat, cnt, launches);
    memset(stats, 0, sizeof(*stats));
    stats->primary_rays = pt.primary;
    stats->reflection_rays = (int64_t)fc.reflections;
    stats->shadow_rays = (int64_t)fc.hits * ds.view.n_lights;
    stats->launches = launches;
    stats->ms_total = cat[CAT_FRAME];
    stats->ms_trace_closest = cat[CAT_CLOSEST];
    stats->ms_trace_any = cat[CAT_ANY];
    stats->ms_shade = cat[CAT_SHADE];
    stats->ms_other = cat[CAT_OTHER];
    stats->max_depth = pt.primary > 0 ? std::max(1, (int)fc.max_depth) : 0;
    stats->truncated_paths = (int64_t)fc.truncated;
    stats->n_gpus = 1;
    stats->n_closest = cnt[CAT_CLOSEST]; stats->n_any = cnt[CAT_ANY]; stats->n_shade = cnt[CAT_SHADE]; stats->n_other = cnt[CAT_OTHER];
    stats->frames = pt.frames;
    pt.frames = 0; pt.primary = 0;
    return YRT_OK;
}

int read_counters_device(DevScene& ds, uint64_t out[24]) {
#if YRT_COUNTERS
    YRT_CUDA(cudaSetDevice(ds.device));
    YRT_CUDA(cudaDeviceSynchronize());
    if (!ds.dctr.p) { memset(out, 0, sizeof(uint64_t) * 24); return YRT_OK; }
    YRT_CUDA(cudaMemcpy(out, ds.dctr.p, sizeof(uint64_t) * 24, cudaMemcpyDeviceToHost));
    YRT_CUDA(cudaMemset(ds.dctr.p, 0, sizeof(uint64_t) * 24));
    return YRT_OK;
#else
    (void)ds; (void)out;
    set_error("this build of the library does not count per-ray work (rebuild with -DYRT_COUNTERS=1, tools/build_variants.sh)");
    return YRT_ERR_UNSUPPORTED;
#endif
}

int read_truncated_device(DevScene& ds, int64_t* out) {
    *out = 0;
    if (!ds.timer || !ds.ws.stats.p) return YRT_OK;   // nothing rendered yet (the counters are cleared at the start of a frame)
    YRT_CUDA(cudaSetDevice(ds.device));
    YRT_CUDA(cudaDeviceSynchronize());
    FrameCounters fc;
    YRT_CUDA(cudaMemcpy(&fc, ds.ws.stats.p, sizeof(fc), cudaMemcpyDeviceToHost));
    *out = (int64_t)fc.truncated;
    return YRT_OK;
}

int stats_begin_device(DevScene& ds) {
    if (!ds.timer) ds.timer = new PhaseTimer();
    PhaseTimer& pt = *ds.timer;
    YRT_CUDA(cudaSetDevice(ds.device));
    YRT_TRY(ds.ws.stats.alloc(STATS_BYTES, ds.device));
    if (pt.main || ds.stream) YRT_CUDA(cudaStreamSynchronize(pt.main ? pt.main : ds.stream));
    YRT_CUDA(cudaMemset(ds.ws.stats.p, 0, sizeof(FrameCounters)));
    for (auto& sp : pt.spans) { pt.pool.push_back(sp.a); pt.pool.push_back(sp.b); }
    pt.spans.clear();
    pt.frames = 0; pt.primary = 0; pt.extra_launches = 0;
    pt.deferred = true;
    return YRT_OK;
}

int stats_end_device(DevScene& ds, yrt_stats* stats) {
    if (!ds.timer || !ds.timer->deferred) { set_error("yrt_stats_end without yrt_stats_begin"); return YRT_ERR_INVALID; }
    RenderParams dummy;
    memset(&dummy, 0, sizeof(dummy));
    ds.timer->deferred = false;
    if (!ds.timer->main) ds.timer->main = ds.stream;
    return collect_stats_device(ds, dummy, stats);
}

int trace_primary_device(DevScene& ds, const RenderParams& rp_in, int32_t* h_ids, float* h_dist, float* h_uv) {
    RenderParams rp = rp_in;
    rp.tile_rows = std::max(rp.height, 1); rp.rank = 0; rp.world = 1;
    YRT_TRY(check_params(rp));
    YRT_CUDA(cudaSetDevice(ds.device));
    cudaStream_t st = ds.stream;
    int spp = rp.samples * rp.samples;
    int batch_rows = batch_rows_for(rp, 1, rp.height, true);
    size_t cap_slots = (size_t)batch_rows * rp.width * spp;
    YRT_TRY(ensure_workspace(ds, ds.ws, cap_slots, ds.view.n_lights, 1, false));
    DevBuf d_ids, d_dist, d_uv;
    YRT_TRY(d_ids.alloc(sizeof(int) * 3 * cap_slots, ds.device));
    YRT_TRY(d_dist.alloc(sizeof(float) * cap_slots, ds.device));
    YRT_TRY(d_uv.alloc(sizeof(float) * 2 * cap_slots, ds.device));
    PhaseTimer pt;
    CounterRing ring;
    YRT_TRY(ring.init(ds.ws, st));
    GridRef cam_grid;
    YRT_TRY(camera_grid_enqueue(ds, rp.cam, rp.width, rp.height, st, &cam_grid));
    for (int lr0 = 0; lr0 < rp.height; lr0 += batch_rows) {
        int nrows = std::min(batch_rows, rp.height - lr0);
        size_t n = (size_t)nrows * rp.width * spp;
        YRT_TRY(run_batch(ds, ds.ws, rp, cam_grid, lr0, nrows, ds.ws.cap_slots, nullptr, st, pt, ring, 1, false, true));
        k_hit_ids<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ds.view, ds.ws.hit.as<float4>(), ds.ws.P.as<float4>(), (int)n, d_ids.as<int>(),
                                                             d_dist.as<float>(), d_uv.as<float>());
        YRT_CUDA(cudaGetLastError());
        size_t off = (size_t)lr0 * rp.width * spp;
        YRT_CUDA(cudaMemcpyAsync(h_ids + 3 * off, d_ids.p, sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, st));
        if (h_dist) YRT_CUDA(cudaMemcpyAsync(h_dist + off, d_dist.p, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
        if (h_uv) YRT_CUDA(cudaMemcpyAsync(h_uv + 2 * off, d_uv.p, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, st));
        YRT_CUDA(cudaStreamSynchronize(st));
    }
    return YRT_OK;
}

int intersect_rays_device(DevScene& ds, const float* h_rays, int64_t n, bool any, int32_t* h_ids, float* h_dist, float* h_uv,
                          uint8_t* h_occ) {
    if (n < 0 || (n > 0 && !h_rays)) { set_error("bad ray array"); return YRT_ERR_INVALID; }
    if (n == 0) return YRT_OK;
    YRT_CUDA(cudaSetDevice(ds.device));
    cudaStream_t st = ds.stream;
    const int64_t chunk = 1 << 22;
    DevBuf ro, rd, hit, P, ids, dist, uv, occ, ctr;
    int64_t c = std::min(chunk, n);
    YRT_TRY(ro.alloc(sizeof(float4) * c, ds.device));
    YRT_TRY(rd.alloc(sizeof(float4) * c, ds.device));
    YRT_TRY(hit.alloc(sizeof(float4) * c, ds.device));
    YRT_TRY(P.alloc(sizeof(float4) * c, ds.device));
    YRT_TRY(ids.alloc(sizeof(int) * 3 * c, ds.device));
    YRT_TRY(dist.alloc(sizeof(float) * c, ds.device));
    YRT_TRY(uv.alloc(sizeof(float) * 2 * c, ds.device));
    YRT_TRY(occ.alloc((size_t)c, ds.device));
    YRT_TRY(ctr.alloc(sizeof(unsigned) * 4, ds.device));
    std::vector<float4> ho(c), hd(c);
    const bool ref = ds.ref.n_nodes > 0;
    int g_c = persistent_grid(ds, ref ? (const void*)k_trace_closest_ref<false> : (const void*)k_trace_closest<false>);
    int g_a = persistent_grid(ds, ref ? (const void*)k_trace_any_rays_ref : (const void*)k_trace_any_rays);
    BatchParams bp;
    memset(&bp, 0, sizeof(bp));
    for (int64_t off = 0; off < n; off += chunk) {
        int64_t m = std::min(chunk, n - off);
        for (int64_t i = 0; i < m; i++) {
            const float* r = h_rays + 8 * (off + i);
            ho[i] = mk4(r[0], r[1], r[2], r[6]);
            hd[i] = mk4(r[3], r[4], r[5], r[7]);
        }
        YRT_CUDA(cudaMemcpyAsync(ro.p, ho.data(), sizeof(float4) * m, cudaMemcpyHostToDevice, st));
        YRT_CUDA(cudaMemcpyAsync(rd.p, hd.data(), sizeof(float4) * m, cudaMemcpyHostToDevice, st));
        YRT_CUDA(cudaMemsetAsync(ctr.p, 0, sizeof(unsigned) * 4, st));
        unsigned need = (unsigned)((m + TRACE_THREADS - 1) / TRACE_THREADS);
        if (any) {
            if (ref) k_trace_any_rays_ref<<<std::max(1u, std::min((unsigned)g_a, need)), TRACE_THREADS, 0, st>>>(ds.view, ds.ref, ro.as<float4>(), rd.as<float4>(), occ.as<uint8_t>(),
                                                                                                          workdist_linear(ctr.as<unsigned>(), (unsigned)m));
            else k_trace_any_rays<<<std::max(1u, std::min((unsigned)g_a, need)), TRACE_THREADS, 0, st>>>(ds.view, ro.as<float4>(), rd.as<float4>(), occ.as<uint8_t>(),
                                                                                                 workdist_linear(ctr.as<unsigned>(), (unsigned)m));
            YRT_CUDA(cudaGetLastError());
            YRT_CUDA(cudaMemcpyAsync(h_occ + off, occ.p, (size_t)m, cudaMemcpyDeviceToHost, st));
        } else {
            if (ref) k_trace_closest_ref<false><<<std::max(1u, std::min((unsigned)g_c, need)), TRACE_THREADS, 0, st>>>(
                ds.view, ds.ref, bp, nullptr, ro.as<float4>(), rd.as<float4>(), hit.as<float4>(), P.as<float4>(), workdist_linear(ctr.as<unsigned>(), (unsigned)m));
            else k_trace_closest<false><<<std::max(1u, std::min((unsigned)g_c, need)), TRACE_THREADS, 0, st>>>(
                ds.view, bp, nullptr, ro.as<float4>(), rd.as<float4>(), hit.as<float4>(), P.as<float4>(), workdist_linear(ctr.as<unsigned>(), (unsigned)m), nullptr);
            k_hit_ids<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(ds.view, hit.as<float4>(), P.as<float4>(), (int)m, ids.as<int>(),
                                                                 dist.as<float>(), uv.as<float>());
            YRT_CUDA(cudaGetLastError());
            YRT_CUDA(cudaMemcpyAsync(h_ids + 3 * off, ids.p, sizeof(int) * 3 * m, cudaMemcpyDeviceToHost, st));
            if (h_dist) YRT_CUDA(cudaMemcpyAsync(h_dist + off, dist.p, sizeof(float) * m, cudaMemcpyDeviceToHost, st));
            if (h_uv) YRT_CUDA(cudaMemcpyAsync(h_uv + 2 * off, uv.p, sizeof(float) * 2 * m, cudaMemcpyDeviceToHost, st));
        }
        YRT_CUDA(cudaStreamSynchronize(st));
    }
    return YRT_OK;
}

void destroy_device_scene(DevScene& ds) {
    cudaSetDevice(ds.device);
    if (ds.stream) {
        cudaStreamSynchronize(ds.stream);
        cudaStreamDestroy(ds.stream);
        ds.stream = nullptr;
    }
    for (int k = 0; k < 3; k++) {
        if (ds.aux_stream[k]) { cudaStreamSynchronize(ds.aux_stream[k]); cudaStreamDestroy(ds.aux_stream[k]); ds.aux_stream[k] = nullptr; }
        if (ds.ev_join[k]) { cudaEventDestroy(ds.ev_join[k]); ds.ev_join[k] = nullptr; }
    }
    if (ds.ev_fork) { cudaEventDestroy(ds.ev_fork); ds.ev_fork = nullptr; }
    if (ds.grid_stream) { cudaStreamSynchronize(ds.grid_stream); cudaStreamDestroy(ds.grid_stream); ds.grid_stream = nullptr; }
    if (ds.ev_grid) { cudaEventDestroy(ds.ev_grid); ds.ev_grid = nullptr; }
    for (int k = 0; k < 4; k++) if (ds.ev_primary_done[k]) { cudaEventDestroy(ds.ev_primary_done[k]); ds.ev_primary_done[k] = nullptr; }
    delete ds.timer;
    ds.timer = nullptr;
    cudaDeviceSynchronize();        // nothing of this scene is in flight any more: its arena may serve the next build
    arena_give_back(ds.arena);
}

int unpack_rows_device(const float4* d_packed, float4* d_full, int width, int height, int tile_rows, int rank, int world, cudaStream_t st) {
    int own = rows_owned(height, tile_rows, rank, world);
    size_t n = (size_t)own * width;
    if (n == 0) return YRT_OK;
    k_unpack_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_packed, d_full, width, own, tile_rows, rank, world);
    YRT_CUDA(cudaGetLastError());
    return YRT_OK;
}

int tonemap_launch(const float4* d_in, uint8_t* d_out, size_t n, cudaStream_t st) {
    if (n == 0) return YRT_OK;
    k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, (uchar4*)d_out, n);
    YRT_CUDA(cudaGetLastError());
    return YRT_OK;
}

int tonemap_device(int device, const float* h_rgba, int width, int height, uint8_t* h_out) {
    YRT_CUDA(cudaSetDevice(device));
    size_t n = (size_t)width * height;
    DevBuf in, out;
    YRT_TRY(in.upload(h_rgba, sizeof(float4) * n, device, 0));
    YRT_TRY(out.alloc(sizeof(uchar4) * n, device));
    k_tonemap<<<(unsigned)((n + 255) / 256), 256>>>(in.as<float4>(), out.as<uchar4>(), n);
    YRT_CUDA(cudaGetLastError());
    YRT_CUDA(cudaMemcpy(h_out, out.p, sizeof(uchar4) * n, cudaMemcpyDeviceToHost));
    return YRT_OK;
}

}  // namespace yrt
