// yrt_work.cuh — work distribution of the persistent traversal kernels.
//
// The reference walks pixels row by row on one thread (src/raytrace.cpp:228-251); any order gives the same image,
// because every camera sample is independent and k_resolve restores the (jj, ii) sum order.  What the order decides
// here is cache locality.  Two schemes:
//
//  * linear (mode 0, the first version): a warp takes the next 32 consecutive slots from one global counter.  At any
//    moment the ~4 700 resident warps of the GPU then cover ~5 full image rows, and the 32 warps of ONE SM are a random
//    sample of that strip: an SM's L1 sees nodes and elements from all over the scene (L1 hit rate 84-87 %, ncu).
//  * SM-affine tiles (mode 1/2): the counter hands out CHUNKS — a tile of tile_w x tile_h pixels with all their samples
//    (mode 1), or a run of `chunk_items` consecutive queue entries (mode 2) — and all warps resident on one SM work
//    through the SAME chunk, 32 slots (one "task") at a time, through a 64-bit state word per SM: the rays in flight
//    on an SM are then neighbours on the screen, walk the same subtrees and share their cache lines.
//
// Per-SM state word: high half = chunk + 1 (0 = nothing fetched yet, 0xffffffff = input exhausted), low half = next
// task of that chunk.  One atomicAdd hands a warp its task; the warp whose ticket equals the task count of the chunk
// fetches the next chunk from the global counter and publishes it with an exchange; tickets beyond that wait for the
// publication (a few hundred ns, once per chunk).  The state is indexed by %smid, but nothing depends on that for
// correctness: warps sharing a state word share a chunk, whoever they are.  The protocol is a template over the
// atomics so that tests/host_emu can run the very same code on host threads (tests/test_host_logic.py).
#pragma once
#include "yrt_math.cuh"

namespace yrt {

#define YRT_SM_STATES 256u          /* state words per launch (>= SM count; %smid is masked into it) */
#define YRT_WORK_BLOCK_WORDS 1024   /* unsigned words of one launch's record: [0] chunk counter, [32..] state words (8-byte aligned) */
#define YRT_STATE_OFFSET_WORDS 32
#define YRT_WORK_BLOCKS 64          /* launch records per workspace ring */

struct WorkDist {
    unsigned* counter;               // next chunk (modes 1, 2) or next item (mode 0)
    unsigned long long* sm_state;    // [YRT_SM_STATES], zeroed before the launch
    int mode;                        // 0 linear per-warp, 1 pixel tiles, 2 runs of queue entries
    int width, nrows, spp;           // mode 1: the batch is nrows x width pixels, spp samples each, slot = pixel * spp + sample
    int tile_w, tile_h, tiles_x;     // mode 1
    unsigned chunk_items;            // mode 2
    unsigned n_items;                // all modes: total work items of the launch
    unsigned n_chunks;               // modes 1, 2
};

// rec = a zeroed record of YRT_WORK_BLOCK_WORDS words (one word is enough for workdist_linear)
inline WorkDist workdist_linear(unsigned* rec, unsigned n) {
    WorkDist wd = {};
    wd.counter = rec;
    wd.sm_state = nullptr;
    wd.mode = 0;
    wd.n_items = n;
    return wd;
}
// the slots of a batch of nrows x width pixels with spp samples each, as tiles of tile_w x tile_h pixels
inline WorkDist workdist_tiles(unsigned* rec, int width, int nrows, int spp, int tile_w, int tile_h) {
    WorkDist wd = {};
    wd.counter = rec;
    wd.sm_state = (unsigned long long*)(rec + YRT_STATE_OFFSET_WORDS);
    wd.mode = 1;
    wd.width = width; wd.nrows = nrows; wd.spp = spp;
    wd.tile_w = tile_w < 1 ? 1 : (tile_w > width ? width : tile_w);
    wd.tile_h = tile_h < 1 ? 1 : (tile_h > nrows ? nrows : tile_h);
    if (wd.tile_w < 1) wd.tile_w = 1;   // empty batch
    if (wd.tile_h < 1) wd.tile_h = 1;
    wd.tiles_x = (width + wd.tile_w - 1) / wd.tile_w;
    wd.n_items = (unsigned)width * (unsigned)nrows * (unsigned)spp;
    wd.n_chunks = (unsigned)wd.tiles_x * (unsigned)((nrows + wd.tile_h - 1) / wd.tile_h);
    return wd;
}
// n queue entries as runs of chunk_items
inline WorkDist workdist_runs(unsigned* rec, unsigned n, unsigned chunk_items) {
    WorkDist wd = {};
    wd.counter = rec;
    wd.sm_state = (unsigned long long*)(rec + YRT_STATE_OFFSET_WORDS);
    wd.mode = 2;
    wd.spp = 1;
    wd.chunk_items = chunk_items < 32u ? 32u : chunk_items;
    wd.n_items = n;
    wd.n_chunks = (n + wd.chunk_items - 1u) / wd.chunk_items;
    return wd;
}

YRT_HD unsigned chunk_items_of(const WorkDist& wd, unsigned chunk, int& x0, int& y0, int& tw, int& th) {
    if (wd.mode == 1) {
        int ty = (int)(chunk / (unsigned)wd.tiles_x), tx = (int)(chunk - (unsigned)ty * (unsigned)wd.tiles_x);
        x0 = tx * wd.tile_w; y0 = ty * wd.tile_h;
        tw = wd.width - x0 < wd.tile_w ? wd.width - x0 : wd.tile_w;
        th = wd.nrows - y0 < wd.tile_h ? wd.nrows - y0 : wd.tile_h;
        return (unsigned)(tw * th) * (unsigned)wd.spp;
    }
    x0 = y0 = 0; tw = th = 1;
    unsigned first = chunk * wd.chunk_items;
    unsigned left = wd.n_items - first;
    return left < wd.chunk_items ? left : wd.chunk_items;
}

YRT_HD unsigned chunk_tasks(const WorkDist& wd, unsigned chunk) {
    int x0, y0, tw, th;
    return (chunk_items_of(wd, chunk, x0, y0, tw, th) + 31u) / 32u;
}

// item (slot of the batch in mode 1, queue index in mode 2) of lane `lane` in task k of `chunk`; false = lane has no item
YRT_HD bool task_item(const WorkDist& wd, unsigned chunk, unsigned k, int lane, unsigned& item) {
    int x0, y0, tw, th;
    unsigned n = chunk_items_of(wd, chunk, x0, y0, tw, th);
    unsigned q = k * 32u + (unsigned)lane;
    if (q >= n) return false;
    if (wd.mode == 1) {
        unsigned p = q / (unsigned)wd.spp, s = q - p * (unsigned)wd.spp;       // pixel of the tile (row-major inside it), sample
        unsigned py = p / (unsigned)tw, px = p - py * (unsigned)tw;
        item = ((unsigned)(y0 + (int)py) * (unsigned)wd.width + (unsigned)(x0 + (int)px)) * (unsigned)wd.spp + s;
    } else {
        item = chunk * wd.chunk_items + q;
    }
    return true;
}

#define YRT_STATE_DONE 0xffffffffu

// One warp's (one host thread's) request for a task.  A: add64/exch64/load64 on the state word, add32 on the counter, pause().
template <class A>
YRT_HD bool fetch_task(A& at, const WorkDist& wd, unsigned sm, unsigned& chunk, unsigned& k) {
    unsigned long long* st = wd.sm_state + (sm & (YRT_SM_STATES - 1u));
    for (;;) {
        unsigned long long old = at.add64(st, 1ull);
        unsigned tag = (unsigned)(old >> 32), ticket = (unsigned)old;
        if (tag == YRT_STATE_DONE) return false;
        unsigned tasks = tag == 0u ? 0u : chunk_tasks(wd, tag - 1u);
        if (ticket < tasks) { chunk = tag - 1u; k = ticket; return true; }
        if (ticket == tasks) {
            // this chunk just ran out (or nothing was fetched yet) and this is the one ticket that says so: fetch the next
            unsigned c = at.add32(wd.counter, 1u);
            if (c >= wd.n_chunks) { at.exch64(st, (unsigned long long)YRT_STATE_DONE << 32); return false; }
            at.exch64(st, ((unsigned long long)(c + 1u) << 32) | 1ull);   // task 0 is ours
            chunk = c; k = 0u;
            return true;
        }
        while ((unsigned)(at.load64(st) >> 32) == tag) at.pause();   // the fetching warp is about to publish
    }
}

#ifdef __CUDACC__
struct DeviceAtomics {
    __device__ __forceinline__ unsigned long long add64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
    __device__ __forceinline__ unsigned long long exch64(unsigned long long* p, unsigned long long v) { return atomicExch(p, v); }
    __device__ __forceinline__ unsigned long long load64(unsigned long long* p) { return *(volatile unsigned long long*)p; }
    __device__ __forceinline__ unsigned add32(unsigned* p, unsigned v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void pause() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 700
        __nanosleep(40);
#endif
    }
};
#endif

}  // namespace yrt
