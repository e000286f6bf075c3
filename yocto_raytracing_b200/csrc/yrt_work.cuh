// yrt_work.cuh — work distribution of the persistent traversal kernels.
//
// The reference walks pixels row by row on one thread (src/raytrace.cpp:228-251); any order gives the same image,
// because every camera sample is independent and k_resolve restores the (jj, ii) sum order.  Here a warp takes the next
// 32 consecutive work items (= the 16 samples of two neighbouring pixels at 16 spp) from one global counter.
// The number of items may live in device memory (n_items_dev): the reflection waves of a frame are launched without the
// host ever reading how many rays the previous wave spawned.
// (Measured alternatives — SM-affine pixel tiles, lane refill, warp packets — are in profiles/r1_experiments.md and in
// the history of this file; none was faster.)
#pragma once
#include "yrt_math.cuh"

namespace yrt {

#define YRT_WORK_BLOCKS 256         /* launch counters per workspace ring */

struct WorkDist {
    unsigned* counter;               // next item; zeroed before the launch
    const unsigned* n_items_dev;     // if not null: the item count is read from here by the kernel
    unsigned n_items;                // item count (upper bound when n_items_dev is set)
};

inline WorkDist workdist_linear(unsigned* counter, unsigned n, const unsigned* n_dev = nullptr) {
    WorkDist wd;
    wd.counter = counter;
    wd.n_items_dev = n_dev;
    wd.n_items = n;
    return wd;
}

}  // namespace yrt
