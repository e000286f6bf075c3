// yrt_api.cu — the extern "C" boundary declared in include/yrt_b200.h.
//
// Multi-GPU inside one process (yrt_render with yrt_init(n>1)): the scene is replicated, GPU g renders the interleaved
// row tiles t with t % n == g and copies its rows ITSELF into the caller's host frame (one pitched device->host copy per
// GPU, each over its own PCIe link, all concurrent) — there is no exchange between GPUs at all.
// One process per GPU (torchrun): yrt_init_device + yrt_render_rows_to_host into a host frame shared by the ranks, or
// yrt_render_rows_into_frame when the frame is wanted in rank 0's device memory (peer stores over NVLink).
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>

#include "yrt_internal.h"

using namespace yrt;

namespace {
std::mutex g_mu;
std::vector<int> g_devices;   // empty until yrt_init*

int g_allow_nonrigid = 0;    // yrt_set_option("allow_nonrigid"): accepted for ABI compatibility, no effect (such scenes are traced through the reference's instance tree)

// With the option "pin_host_frames" the caller's frame buffers are page-locked on first sight and stay so until another
// buffer takes the slot or the library is re-initialised (device->host copies into pageable memory are staged by the
// driver).  Opt-in, because the registration outlives the call: the caller must keep such a buffer alive (and not hand its
// address range back to the allocator) while it is registered.
struct HostPin { void* p = nullptr; size_t n = 0; };
HostPin g_pins[2];
int g_pin_host_frames = 0;   // yrt_set_option("pin_host_frames")
void pin_host(int slot, void* p, size_t n) {
    if (!g_pin_host_frames || n < ((size_t)1 << 20)) return;
    HostPin& h = g_pins[slot];
    if (h.p == p && h.n == n) return;
    if (h.p) { cudaHostUnregister(h.p); h.p = nullptr; }
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) { cudaGetLastError(); return; }   // pinned already
    cudaGetLastError();
    if (cudaHostRegister(p, n, cudaHostRegisterPortable) == cudaSuccess) { h.p = p; h.n = n; }
    cudaGetLastError();
}
void unpin_all() {
    for (HostPin& h : g_pins)
        if (h.p) { cudaHostUnregister(h.p); h.p = nullptr; h.n = 0; }
    cudaGetLastError();
}

int ensure_init() {
    if (!g_devices.empty()) return YRT_OK;
    return yrt_init(1);
}
}  // namespace

extern "C" {

int yrt_abi_version(void) { return YRT_ABI_VERSION; }
const char* yrt_last_error(void) { return get_error(); }

int yrt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int yrt_init(int n_gpus) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (n_gpus <= 0) n_gpus = 1;
    int have = yrt_device_count();
    if (have < n_gpus) {
        set_error("yrt_init: %d CUDA device(s) requested, %d visible — this library has no CPU path", n_gpus, have);
        return YRT_ERR_NO_DEVICE;
    }
    g_devices.clear();
    for (int i = 0; i < n_gpus; i++) g_devices.push_back(i);
    for (int i = n_gpus - 1; i >= 0; i--) {   // create the contexts now, not inside the first scene build
        YRT_CUDA(cudaSetDevice(i));
        YRT_CUDA(cudaFree(0));
    }
    unpin_all();
    return YRT_OK;
}

int yrt_init_device(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    int have = yrt_device_count();
    if (device < 0 || device >= have) {
        set_error("yrt_init_device: device %d not available (%d visible) — this library has no CPU path", device, have);
        return YRT_ERR_NO_DEVICE;
    }
    g_devices.assign(1, device);
    YRT_CUDA(cudaSetDevice(device));
    YRT_CUDA(cudaFree(0));
    unpin_all();
    return YRT_OK;
}

int yrt_scene_create(const yrt_scene_desc* desc, yrt_scene** out) {
    if (!out) { set_error("yrt_scene_create: out is null"); return YRT_ERR_INVALID; }
    *out = nullptr;
    yrt_scene* s = new yrt_scene();
    int st = host_scene_from_desc(desc, s->host, true);
    if (st != YRT_OK) { delete s; return st; }
    st = ensure_init();
    if (st != YRT_OK) { delete s; return st; }
    for (int dev : g_devices) {
        DevScene* ds = new DevScene();
        s->dev.push_back(ds);
        st = build_device_scene(s->host, dev, *ds);
        if (st != YRT_OK) { yrt_scene_destroy(s); return st; }
    }
    *out = s;
    return YRT_OK;
}

void yrt_scene_destroy(yrt_scene* scn) {
    if (!scn) return;
    for (DevScene* ds : scn->dev) {
        destroy_device_scene(*ds);
        delete ds;
    }
    delete scn;
}

int yrt_scene_info(const yrt_scene* scn, int64_t out[8]) {
    if (!scn || scn->dev.empty() || !out) { set_error("yrt_scene_info: bad arguments"); return YRT_ERR_INVALID; }
    const DevScene& d = *scn->dev[0];
    out[0] = d.n_blas_nodes; out[1] = d.n_tlas_nodes; out[2] = d.blas_depth; out[3] = d.tlas_depth;
    out[4] = d.view.n_lights; out[5] = d.n_prims; out[6] = (int64_t)d.build_us; out[7] = scn->host.n_reflective;
    return YRT_OK;
}

int yrt_desc_nonrigid_instances(const yrt_scene_desc* desc) {
    if (!desc || (desc->n_instances > 0 && !desc->inst_frame)) return 0;
    int n = 0;
    for (int i = 0; i < desc->n_instances; i++) {
        const float* f = desc->inst_frame + 12 * (size_t)i;
        auto dot3 = [](const float* a, const float* b) { return (double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2]; };
        const double tol = 1e-4;
        bool rigid = fabs(dot3(f, f) - 1.0) <= tol && fabs(dot3(f + 3, f + 3) - 1.0) <= tol && fabs(dot3(f + 6, f + 6) - 1.0) <= tol &&
                     fabs(dot3(f, f + 3)) <= tol && fabs(dot3(f, f + 6)) <= tol && fabs(dot3(f + 3, f + 6)) <= tol;
        if (!rigid) n++;
    }
    return n;
}

int yrt_set_option(const char* name, int value) {
    if (name && !strcmp(name, "allow_nonrigid")) { g_allow_nonrigid = value != 0; return YRT_OK; }
    if (name && !strcmp(name, "pin_host_frames")) { g_pin_host_frames = value != 0; if (!value) unpin_all(); return YRT_OK; }
    set_error("yrt_set_option: unknown option '%s'", name ? name : "(null)");
    return YRT_ERR_INVALID;
}

int yrt_scene_prepare(yrt_scene* scn, int width, int height, int samples) {
    if (!scn || scn->dev.empty() || width <= 0 || height <= 0 || samples <= 0) { set_error("yrt_scene_prepare: bad arguments"); return YRT_ERR_INVALID; }
    const int G = (int)scn->dev.size();
    for (DevScene* ds : scn->dev) YRT_TRY(presize_workspace_device(*ds, width, (height + G - 1) / G, samples));
    return YRT_OK;
}

int64_t yrt_debug_read_nodes(yrt_scene* scn, int arity, float* out, int64_t max_float4) {
    if (!scn || scn->dev.empty() || (arity != 2 && arity != 4)) { set_error("yrt_debug_read_nodes: bad arguments"); return YRT_ERR_INVALID; }
    DevScene& ds = *scn->dev[0];
    const int64_t n = (int64_t)YRT_NODE_STRIDE(arity) * (ds.n_blas_nodes + ds.n_tlas_nodes);
    if (!out) return n;
    if (max_float4 < n) { set_error("yrt_debug_read_nodes: buffer too small"); return YRT_ERR_INVALID; }
    YRT_CUDA(cudaSetDevice(ds.device));
    YRT_CUDA(cudaMemcpy(out, arity == 4 ? ds.view.nodes4 : ds.view.nodes2, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost));
    return n;
}

int yrt_counters_read(yrt_scene* scn, uint64_t out[24]) {
    if (!scn || scn->dev.empty() || !out) { set_error("yrt_counters_read: bad arguments"); return YRT_ERR_INVALID; }
    return read_counters_device(*scn->dev[0], out);
}

int yrt_frame_truncated_paths(yrt_scene* scn, int64_t* out) {
    if (!scn || scn->dev.empty() || !out) { set_error("yrt_frame_truncated_paths: bad arguments"); return YRT_ERR_INVALID; }
    int64_t total = 0;
    for (DevScene* ds : scn->dev) {
        int64_t n = 0;
        YRT_TRY(read_truncated_device(*ds, &n));
        total += n;
    }
    *out = total;
    return YRT_OK;
}

int yrt_write_png(const char* path, const uint8_t* rgba8, int width, int height, int threads, int level) {
    return write_png_parallel(path, rgba8, width, height, threads, level <= 0 ? 1 : (level > 9 ? 9 : level));
}

int yrt_image_width(const yrt_camera* cam, int resolution) {
    if (!cam) return 0;
    return (int)roundf(cam->aspect * (float)resolution);   // (int)std::round(cam->aspect * resolution)
}

int yrt_rows_owned(int height, int tile_rows, int rank, int world) {
    if (height <= 0 || tile_rows <= 0 || world <= 0 || rank < 0 || rank >= world) return 0;
    return rows_owned(height, tile_rows, rank, world);
}

static int fill_params(const yrt_camera* cam, const float amb[3], int width, int height, int samples, RenderParams& rp) {
    if (!cam) { set_error("camera is null"); return YRT_ERR_INVALID; }
    rp.cam = make_camera_k(cam);
    rp.amb = amb ? mk3(amb[0], amb[1], amb[2]) : mk3(0.f, 0.f, 0.f);
    rp.width = width; rp.height = height; rp.samples = samples;
    rp.tile_rows = 1; rp.rank = 0; rp.world = 1;
    return YRT_OK;
}

int yrt_render_rows(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples, int tile_rows,
                    int rank, int world, void* d_rgba, void* stream, yrt_stats* stats) {
    if (!scn || scn->dev.empty() || !d_rgba) { set_error("yrt_render_rows: bad arguments"); return YRT_ERR_INVALID; }
    RenderParams rp;
    YRT_TRY(fill_params(cam, amb, width, height, samples, rp));
    rp.tile_rows = tile_rows; rp.rank = rank; rp.world = world;
    DevScene& ds = *scn->dev[0];
    cudaStream_t st = (cudaStream_t)stream;   // NULL = legacy default stream
    return render_rows_device(ds, rp, (float4*)d_rgba, st, stats, true);
}

int yrt_render_rows_into_frame(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples,
                               int tile_rows, int rank, int world, void* d_full, void* stream, yrt_stats* stats) {
    if (!scn || scn->dev.empty() || !d_full) { set_error("yrt_render_rows_into_frame: bad arguments"); return YRT_ERR_INVALID; }
    RenderParams rp;
    YRT_TRY(fill_params(cam, amb, width, height, samples, rp));
    rp.tile_rows = tile_rows; rp.rank = rank; rp.world = world;
    rp.scatter = true;
    return render_rows_device(*scn->dev[0], rp, (float4*)d_full, (cudaStream_t)stream, stats, true);
}

int yrt_host_barrier(void* counter, int world, int64_t generation) {
    if (!counter || world <= 0 || generation <= 0) { set_error("yrt_host_barrier: bad arguments"); return YRT_ERR_INVALID; }
    volatile int64_t* c = (volatile int64_t*)counter;
    __atomic_fetch_add((int64_t*)counter, (int64_t)1, __ATOMIC_ACQ_REL);
    const int64_t target = (int64_t)world * generation;
    for (unsigned spins = 0; __atomic_load_n((int64_t*)c, __ATOMIC_ACQUIRE) < target; spins++)
        if (spins > 2000) std::this_thread::yield();
    return YRT_OK;
}

// The shared frame allocation holds TWO frames back to back and, behind them, the arrival counter of yrt_frame_barrier
// (256 bytes, zeroed): [frame 0 | frame 1 | counter].
static size_t frame_bytes(int width, int height) { return sizeof(float4) * (size_t)width * height; }

int yrt_frame_alloc(int width, int height, void** d_full) {
    if (!d_full || width <= 0 || height <= 0) { set_error("yrt_frame_alloc: bad arguments"); return YRT_ERR_INVALID; }
    YRT_TRY(ensure_init());
    YRT_CUDA(cudaSetDevice(g_devices[0]));
    YRT_CUDA(cudaMalloc(d_full, 2 * frame_bytes(width, height) + 256));
    YRT_CUDA(cudaMemset((char*)*d_full + 2 * frame_bytes(width, height), 0, 256));
    return YRT_OK;
}

// One thread per rank: this rank's arrival, then wait until `target` arrivals have been counted.  The counter lives in rank
// 0's HBM; the other ranks reach it through their CUDA-IPC mapping of the frame (NVLink peer atomics and loads).  The
// stores of this rank's rows (earlier kernels of the same stream, into the same peer memory) are fenced before the arrival.
__global__ void k_frame_barrier(unsigned long long* ctr, unsigned long long target) {
    __threadfence_system();
    atomicAdd_system(ctr, 1ull);
    while (*(volatile unsigned long long*)ctr < target) __nanosleep(100);
    __threadfence_system();
}

int yrt_frame_barrier(void* d_full, int width, int height, int world, int64_t generation, void* stream) {
    if (!d_full || width <= 0 || height <= 0 || world <= 0 || generation <= 0) { set_error("yrt_frame_barrier: bad arguments"); return YRT_ERR_INVALID; }
    unsigned long long* ctr = (unsigned long long*)((char*)d_full + 2 * frame_bytes(width, height));
    k_frame_barrier<<<1, 1, 0, (cudaStream_t)stream>>>(ctr, (unsigned long long)world * (unsigned long long)generation);
    YRT_CUDA(cudaGetLastError());
    return YRT_OK;
}

int yrt_frame_free(void* d_full) {
    if (d_full) YRT_CUDA(cudaFree(d_full));
    return YRT_OK;
}

int yrt_frame_export(void* d_full, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_full || !handle) { set_error("yrt_frame_export: bad arguments"); return YRT_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    YRT_CUDA(cudaIpcGetMemHandle(&h, d_full));
    memcpy(handle, &h, 64);
    return YRT_OK;
}

int yrt_frame_import(const unsigned char handle[64], void** d_full) {
    if (!d_full || !handle) { set_error("yrt_frame_import: bad arguments"); return YRT_ERR_INVALID; }
    YRT_TRY(ensure_init());
    YRT_CUDA(cudaSetDevice(g_devices[0]));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    YRT_CUDA(cudaIpcOpenMemHandle(d_full, h, cudaIpcMemLazyEnablePeerAccess));
    return YRT_OK;
}

int yrt_frame_release(void* d_full) {
    if (d_full) YRT_CUDA(cudaIpcCloseMemHandle(d_full));
    return YRT_OK;
}

int yrt_stats_begin(yrt_scene* scn) {
    if (!scn || scn->dev.empty()) { set_error("yrt_stats_begin: bad arguments"); return YRT_ERR_INVALID; }
    return stats_begin_device(*scn->dev[0]);
}

int yrt_stats_end(yrt_scene* scn, yrt_stats* totals) {
    if (!scn || scn->dev.empty() || !totals) { set_error("yrt_stats_end: bad arguments"); return YRT_ERR_INVALID; }
    return stats_end_device(*scn->dev[0], totals);
}

int yrt_unpack_rows(const void* d_packed, void* d_full, int width, int height, int tile_rows, int rank, int world, void* stream) {
    if (!d_packed || !d_full || width <= 0 || height <= 0 || tile_rows <= 0 || world <= 0 || rank < 0 || rank >= world) {
        set_error("yrt_unpack_rows: bad arguments");
        return YRT_ERR_INVALID;
    }
    return unpack_rows_device((const float4*)d_packed, (float4*)d_full, width, height, tile_rows, rank, world, (cudaStream_t)stream);
}

int yrt_render_rows_to_host(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples, int tile_rows,
                            int rank, int world, float* h_frame, void* stream, yrt_stats* stats) {
    if (!scn || scn->dev.empty() || !h_frame) { set_error("yrt_render_rows_to_host: bad arguments"); return YRT_ERR_INVALID; }
    RenderParams rp;
    YRT_TRY(fill_params(cam, amb, width, height, samples, rp));
    rp.tile_rows = tile_rows; rp.rank = rank; rp.world = world;
    DevScene& ds = *scn->dev[0];
    const int own = rows_owned(height, tile_rows, rank, world);
    YRT_TRY(ds.ws.rows.alloc(sizeof(float4) * (size_t)std::max(own, 1) * width, ds.device));
    cudaStream_t st = (cudaStream_t)stream;
    rp.h_rgba = h_frame;
    YRT_TRY(render_rows_device(ds, rp, ds.ws.rows.as<float4>(), st, stats, false));
    if (stats) { YRT_CUDA(cudaStreamSynchronize(st)); return collect_stats_device(ds, rp, stats); }
    return YRT_OK;
}

static int render_impl(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples, float* rgba_out,
                       uint8_t* ldr_out, yrt_stats* stats) {
    if (!scn || scn->dev.empty() || (!rgba_out && !ldr_out)) { set_error("yrt_render: bad arguments"); return YRT_ERR_INVALID; }
    RenderParams rp0;
    YRT_TRY(fill_params(cam, amb, width, height, samples, rp0));
    const int G = (int)scn->dev.size();
    const char* etr = getenv("YRT_TILE_ROWS");
    int tile_rows = etr ? std::max(1, atoi(etr)) : 1;   // 1-row tiles: rows r, r+G, r+2G, ... balance best (cost varies smoothly down the image)
    if (G == 1) tile_rows = std::max(1, height);
    const size_t npx = (size_t)width * height;
    if (rgba_out) pin_host(0, rgba_out, sizeof(float4) * npx);
    if (ldr_out) pin_host(1, ldr_out, 4 * npx);

    // every GPU renders its interleaved rows and copies them itself into the caller's frame(s), over its own PCIe link
    std::vector<int> status(G, YRT_OK);
    std::vector<std::string> errs(G);
    auto work = [&](int g) {
        DevScene& ds = *scn->dev[g];
        RenderParams rp = rp0;
        rp.tile_rows = tile_rows; rp.rank = g; rp.world = G;
        const int own = rows_owned(height, tile_rows, g, G);
        int s = ds.ws.rows.alloc(sizeof(float4) * (size_t)std::max(own, 1) * width, ds.device);
        rp.h_rgba = rgba_out;
        if (s == YRT_OK && ldr_out) {   // tonemap on the device: a quarter of the bytes cross to the host
            s = ds.ws.rows8.alloc(4 * (size_t)std::max(own, 1) * width, ds.device);
            rp.h_ldr = ldr_out;
            rp.d_ldr_rows = ds.ws.rows8.as<uint8_t>();
        }
        if (s == YRT_OK) s = render_rows_device(ds, rp, ds.ws.rows.as<float4>(), ds.stream, stats, false);
        if (s == YRT_OK && cudaStreamSynchronize(ds.stream) != cudaSuccess) { set_error("device %d: %s", ds.device, cudaGetErrorString(cudaGetLastError())); s = YRT_ERR_CUDA; }
        status[g] = s;
        if (s != YRT_OK) errs[g] = get_error();
    };
    if (G == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < G; g++)
        if (status[g] != YRT_OK) { set_error("%s", errs[g].c_str()); return status[g]; }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (int g = 0; g < G; g++) {
            DevScene& ds = *scn->dev[g];
            RenderParams rp = rp0;
            rp.tile_rows = tile_rows; rp.rank = g; rp.world = G;
            yrt_stats one;
            YRT_TRY(collect_stats_device(ds, rp, &one));
            stats->primary_rays += one.primary_rays;
            stats->reflection_rays += one.reflection_rays;
            stats->shadow_rays += one.shadow_rays;
            stats->truncated_paths += one.truncated_paths;
            stats->launches += one.launches;
            stats->ms_total = std::max(stats->ms_total, one.ms_total);
            stats->max_depth = std::max(stats->max_depth, one.max_depth);
            if (g == 0) {
                stats->ms_trace_closest = one.ms_trace_closest; stats->ms_trace_any = one.ms_trace_any;
                stats->ms_shade = one.ms_shade; stats->ms_other = one.ms_other;
                stats->n_closest = one.n_closest; stats->n_any = one.n_any; stats->n_shade = one.n_shade; stats->n_other = one.n_other;
            }
        }
        stats->n_gpus = G;
    }
    return YRT_OK;
}

int yrt_render(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples, float* rgba_out,
               yrt_stats* stats) {
    if (!rgba_out) { set_error("yrt_render: rgba_out is null"); return YRT_ERR_INVALID; }
    return render_impl(scn, cam, amb, width, height, samples, rgba_out, nullptr, stats);
}

int yrt_render_ldr(yrt_scene* scn, const yrt_camera* cam, const float amb[3], int width, int height, int samples, uint8_t* rgba8_out,
                   float* rgba_out, yrt_stats* stats) {
    if (!rgba8_out) { set_error("yrt_render_ldr: rgba8_out is null"); return YRT_ERR_INVALID; }
    return render_impl(scn, cam, amb, width, height, samples, rgba_out, rgba8_out, stats);
}

int yrt_trace_primary(yrt_scene* scn, const yrt_camera* cam, int width, int height, int samples, int32_t* ids_out, float* dist_out,
                      float* uv_out) {
    if (!scn || scn->dev.empty() || !ids_out) { set_error("yrt_trace_primary: bad arguments"); return YRT_ERR_INVALID; }
    RenderParams rp;
    float amb[3] = {0, 0, 0};
    YRT_TRY(fill_params(cam, amb, width, height, samples, rp));
    return trace_primary_device(*scn->dev[0], rp, ids_out, dist_out, uv_out);
}

int yrt_intersect_first(yrt_scene* scn, const float* rays, int64_t n, int32_t* ids_out, float* dist_out, float* uv_out) {
    if (!scn || scn->dev.empty() || (n > 0 && !ids_out)) { set_error("yrt_intersect_first: bad arguments"); return YRT_ERR_INVALID; }
    return intersect_rays_device(*scn->dev[0], rays, n, false, ids_out, dist_out, uv_out, nullptr);
}

int yrt_intersect_any(yrt_scene* scn, const float* rays, int64_t n, uint8_t* occluded_out) {
    if (!scn || scn->dev.empty() || (n > 0 && !occluded_out)) { set_error("yrt_intersect_any: bad arguments"); return YRT_ERR_INVALID; }
    return intersect_rays_device(*scn->dev[0], rays, n, true, nullptr, nullptr, nullptr, occluded_out);
}

int yrt_tonemap(const float* rgba_in, int width, int height, uint8_t* rgba8_out) {
    if (!rgba_in || !rgba8_out || width <= 0 || height <= 0) { set_error("yrt_tonemap: bad arguments"); return YRT_ERR_INVALID; }
    YRT_TRY(ensure_init());
    return tonemap_device(g_devices[0], rgba_in, width, height, rgba8_out);
}

}  // extern "C"
