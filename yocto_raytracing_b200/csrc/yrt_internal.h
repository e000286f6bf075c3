// yrt_internal.h — host-side plumbing shared by the translation units of libyrt_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <thread>
#include <vector>

#include "../../include/yrt_b200.h"
#include "yrt_lbvh.cuh"
#include "yrt_pgrid.cuh"
#include "yrt_scene.cuh"

namespace yrt {

void set_error(const char* fmt, ...);
const char* get_error();

#define YRT_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            ::yrt::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorMemoryAllocation) ? YRT_ERR_OOM : YRT_ERR_CUDA;                  \
        }                                                                                           \
    } while (0)

#define YRT_TRY(expr)              \
    do {                           \
        int s_ = (expr);           \
        if (s_ != YRT_OK) return s_; \
    } while (0)

// device buffer owned by a DevScene / workspace (freed on the device it was allocated on)
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int device = -1;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t n, int dev);            // (re)allocates only when growing
    int upload(const void* src, size_t n, int dev, cudaStream_t st);
    void release();
    template <class T> T* as() const { return (T*)p; }
};

// validated host copy of the scene description + everything derived on the host
struct HostScene {
    int n_shapes = 0, n_instances = 0, n_materials = 0, n_textures = 0, n_verts = 0, n_prims = 0;
    std::vector<int> shape_kind, shape_elem_off, shape_elem_cnt, shape_vert_off, shape_vert_cnt, shape_prim_off;
    std::vector<int> shape_has_uv;
    std::vector<int> elem_idx, prim_shape;
    std::vector<float> pos, norm, uv, radius;
    std::vector<float> inst_frame;
    std::vector<int> inst_shape, inst_mat;
    std::vector<int> active_inst;            // instances whose shape has at least one element
    std::vector<int> prim_rank;              // [n_prims] rank of each element in the reference's shape-BVH visit order
    std::vector<int> inst_rank;              // [n_instances] rank of each instance in the reference's scene-BVH visit order
    std::vector<float4> mat_recs;            // 4 per material (ns computed with the host libm)
    std::vector<float4> light_recs;          // 5 per light, in instance order
    std::vector<int> light_inst;
    std::vector<int4> tex_info;
    std::vector<uint8_t> tex_rgba8;
    float srgb_lut[256];
    int n_reflective = 0;
    float extent = 0.f;                      // bound of |coordinate| over everything in the scene (instances' world boxes)
    bool all_rigid = true;                   // every instance frame is orthonormal (to 1e-4): a local hit distance is the world distance
    std::vector<float4> ref_nodes;           // !all_rigid: the reference's instance tree (RefTlas, yrt_scene.cuh), 2 quads per node ...
    std::vector<int> ref_leaf_inst;          // ... and its leaf_prims (instance ids in partition order); filled by reference_visit_ranks
    // the rank tables are computed on a worker thread that yrt_scene_create starts; the build waits for it at the end
    std::thread rank_thread;
    void wait_ranks() { if (rank_thread.joinable()) rank_thread.join(); }
    ~HostScene() { wait_ranks(); }
};

// validates and copies the description; ranks_async: reference_visit_ranks runs on hs.rank_thread instead of inline
int host_scene_from_desc(const yrt_scene_desc* d, HostScene& hs, bool ranks_async = false);
// visit ranks of the reference's top-down BVHs (make_node/split_prims, src/scene.cpp:572-639), for tie-breaking
void reference_visit_ranks(HostScene& hs);

struct PhaseTimer;

// per-device render workspace (grown on demand, reused across frames)
struct Workspace {
    DevBuf hit, P, vis, rad, ray_o, ray_d, act0, act1, counters, stats, rows, rows8;
    DevBuf pstack_lvl[66], pstack_tab;   // per recursion level: {c, kr, la} per slot, allocated when first reached; table of their pointers
    float4** h_pstack_tab = nullptr;     // pinned host copy of the table
    int pstack_levels = 0;
    size_t cap_slots = 0;
    int cap_lights = 0;
    int cap_depth = 0;
    unsigned* h_counts = nullptr;        // pinned: per-wave active-sample counts as they arrive (reflective scenes)
    cudaEvent_t ev_counts[66] = {};      // one per wave (YRT_MAX_WAVES + 2)
    ~Workspace() {
        if (h_counts) { cudaFreeHost(h_counts); for (auto e : ev_counts) if (e) cudaEventDestroy(e); }
        if (h_pstack_tab) cudaFreeHost(h_pstack_tab);
    }
};

struct DevScene {
    int device = 0;
    SceneView view;
    RefTlas ref = {nullptr, nullptr, nullptr, 0};   // n_nodes > 0: the scene has non-rigid instance frames and is traced through the reference's instance tree
    DevBuf arena;          // everything a frame reads: nodes, element / instance / material / light records, textures, rank tables
    int n_prims = 0, n_active = 0, n_blas_nodes = 0, n_tlas_nodes = 0;
    int blas_depth = 0, tlas_depth = 0, stack_need = 0;
    float build_us = 0.f;
    cudaStream_t stream = nullptr;   // owned
    Workspace ws;
    Workspace ws_aux[3];             // further pipelines (part frames on several streams overlap kernel tails)
    cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};   // owned
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    int sm_count = 148;
    bool has_reflective = false;
    int grid_closest_primary = 0, grid_closest_queue = 0, grid_any = 0;   // persistent grids (SMs x resident CTAs)
    DevBuf dctr;                     // per-ray work counters of the traversal kernels (-DYRT_COUNTERS=1 builds)
    // apex grids (yrt_pgrid.cuh): one cube grid per point light, built with the scene (in `arena`); the camera grid is
    // rebuilt by every render call into cg_* (grown on demand)
    LightGrids light_grids;
    int light_grid_R = 0;
    bool grids_allowed = false;      // rigid frames only: the lists are sorted and cut by world-space distances
    float extent = 0.f;
    DevBuf cg_cells, cg_entries, cg_keys, cg_cnt;   // cg_cnt: [n_cells] counts | total | big list (1 + n_active)
    // The camera grid is built on its own stream: its only readers are the primary-ray kernels, so the build of frame k + 1
    // may run as soon as those of frame k are done — under frame k's shadow / shade kernels when frames are enqueued back to back
    cudaStream_t grid_stream = nullptr;             // owned
    cudaEvent_t ev_grid = nullptr, ev_primary_done[4] = {nullptr, nullptr, nullptr, nullptr};
    bool primary_recorded[4] = {false, false, false, false};
    PhaseTimer* timer = nullptr;     // owned (yrt_render.cu)
};

int build_device_scene(HostScene& hs, int device, DevScene& ds);
// enqueues the build of one apex grid on `st` (no host synchronisation): a.cnt / a.total / a.big are cleared first
int pgrid_build_enqueue(const PGridArrays& a, cudaStream_t st);
// this frame's camera grid (ds.cg_*): enqueued on `st`; `out` is what the primary-ray kernel reads (nx = 0: none)
int camera_grid_enqueue(DevScene& ds, const camera_k& cam, int width, int height, cudaStream_t st, GridRef* out);
void destroy_device_scene(DevScene& ds);
void arena_give_back(DevBuf& arena);   // a destroyed scene's arena goes to a small per-device pool for the next build (yrt_build.cu)

struct RenderParams {
    camera_k cam;
    vec3 amb;
    int width, height, samples;
    int tile_rows, rank, world;   // interleaved row tiles
    bool scatter = false;         // d_out is the full frame: rows go to their final position (fused gather)
    // optional host destinations (whole-image, row-major): every batch's rows are copied to their final positions right
    // behind its resolve, on the batch's own stream — the copy of one batch runs under the kernels of the next
    float* h_rgba = nullptr;      // width*height*4 floats
    uint8_t* h_ldr = nullptr;     // width*height*4 bytes: tonemap(hdr, 0, false) on the device first (needs d_ldr_rows)
    uint8_t* d_ldr_rows = nullptr;   // device scratch for the rank's tonemapped packed rows
};

YRT_HD int rows_owned(int height, int tile_rows, int rank, int world) {
    int n_tiles = (height + tile_rows - 1) / tile_rows;
    int rows = 0;
    for (int t = rank; t < n_tiles; t += world) {
        int r0 = t * tile_rows, r1 = r0 + tile_rows;
        if (r1 > height) r1 = height;
        rows += r1 - r0;
    }
    return rows;
}
// global row of the rank's packed local row lr
YRT_HD int global_row(int lr, int tile_rows, int rank, int world) {
    int tl = lr / tile_rows;
    return (tl * world + rank) * tile_rows + (lr - tl * tile_rows);
}

// renders the rank's rows into d_out (packed rows, float4 per pixel) on ds.device / `st`.
// ids/dist/uvw (device, per primary ray) are filled instead of shading when ids != nullptr.
int render_rows_device(DevScene& ds, const RenderParams& rp, float4* d_out, cudaStream_t st, yrt_stats* stats,
                       bool sync_for_stats);
// waits for the frame issued by render_rows_device(…, stats != null, sync_for_stats = false) and fills stats
int collect_stats_device(DevScene& ds, const RenderParams& rp, yrt_stats* stats);
// allocates the render workspace of a (width x height, samples) frame now instead of inside the first render call
int presize_workspace_device(DevScene& ds, int width, int height, int samples);
// per-ray work counters (-DYRT_COUNTERS=1 builds): 3 kernel classes x 8 words, read and reset
int read_counters_device(DevScene& ds, uint64_t out[24]);
// mirror bounces the last stand-alone frame of this device dropped at the recursion cap (waits for the device)
int read_truncated_device(DevScene& ds, int64_t* out);
int stats_begin_device(DevScene& ds);
int stats_end_device(DevScene& ds, yrt_stats* stats);
int trace_primary_device(DevScene& ds, const RenderParams& rp, int32_t* h_ids, float* h_dist, float* h_uv);
int intersect_rays_device(DevScene& ds, const float* h_rays, int64_t n, bool any, int32_t* h_ids, float* h_dist,
                          float* h_uv, uint8_t* h_occ);
int unpack_rows_device(const float4* d_packed, float4* d_full, int width, int height, int tile_rows, int rank,
                       int world, cudaStream_t st);
int tonemap_launch(const float4* d_in, uint8_t* d_out, size_t n, cudaStream_t st);
// packed local rows [lr0, lr0 + nrows) of a rank -> their places in a row-major HOST frame (elem bytes per pixel), on `st`
int copy_rows_to_host(const void* d_packed, void* h_frame, int width, int height, int tile_rows, int rank, int world, size_t elem, int lr0, int nrows,
                      cudaStream_t st);
int tonemap_device(int device, const float* h_rgba, int width, int height, uint8_t* h_out);

camera_k make_camera_k(const yrt_camera* cam);
int write_png_parallel(const char* path, const uint8_t* rgba, int width, int height, int threads, int level);   // yrt_png.cu (host only)

}  // namespace yrt

struct yrt_scene {
    yrt::HostScene host;
    std::vector<yrt::DevScene*> dev;   // one per initialised GPU
};
