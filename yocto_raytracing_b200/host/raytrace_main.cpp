// raytrace_main.cpp — the drop-in `raytrace` CLI: the reference's main() (src/raytrace.cpp:256-287)
// with build_bvh(scn,false) (:278) and raytrace(scn, amb, resolution, samples) (:282) replaced by the
// C ABI of include/yrt_b200.h.  Everything else is the reference's own code, linked unchanged:
// load_scene (src/scene.cpp:113), save_hdr_or_ldr / save_image (src/image.cpp:81,:36), yu::cmdline
// (src/ext/yocto_utils.h:1085+).  Same flags, same four progress lines; additive flags only:
//   --gpus N       interleaved row tiles over N GPUs
//   --cache        keep the flattened scene next to the OBJ (<scene>.yrts) and reuse it while it is not older than
//                  the scene file, any material library / buffer file next to it or any texture it uses: skips load_scene (SURVEY 8f.2 — the OBJ parse costs more than a frame)
//   --device-ldr   tonemap on the GPU (yrt_render_ldr): a quarter of the bytes cross to the host and
//                  the host tonemap (src/image.cpp:55-78) is skipped; ignored for .hdr outputs
//   --fast-png     8-bit PNG outputs are written by yrt_write_png (parallel deflate) instead of stb_image_write: same pixels,
//                  a fraction of the time (stbi_write_png costs 45 frames' worth of rendering at 1080p)
//   --stats        ray counts and the time of every phase
//   --allow-nonrigid  accepted and ignored (earlier versions refused scenes with scaled / sheared instance frames without it; they are
//                  now traced through a copy of the reference's own instance tree and match it, include/yrt_b200.h)
#include <dirent.h>
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <cctype>
#include <cstdlib>

#include "ext/yocto_utils.h"   // reference
#include "scene.h"             // reference
#include "yrt_flatten.h"

namespace {
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// The cache is valid if it exists and is not older than the scene file, nor than any material library / buffer file next
// to it (the loader resolves mtllib and glTF buffer URIs relative to the scene's directory; which of them a scene uses is
// only known after parsing, so all count), nor — checked after loading it — than any texture file it lists.
std::string dir_of(const std::string& path) {
    auto slash = path.find_last_of('/');
    return slash == std::string::npos ? std::string(".") : path.substr(0, slash);
}
bool cache_is_fresh(const std::string& scene_path, const std::string& cache_path) {
    struct stat a, b;
    if (stat(scene_path.c_str(), &a) != 0 || stat(cache_path.c_str(), &b) != 0) return false;
    if (b.st_mtime < a.st_mtime) return false;
    const std::string dir = dir_of(scene_path);
    DIR* d = opendir(dir.c_str());
    if (!d) return true;
    bool fresh = true;
    while (dirent* e = readdir(d)) {
        std::string name = e->d_name;
        auto dot = name.find_last_of('.');
        if (dot == std::string::npos) continue;
        std::string ext = name.substr(dot);
        for (auto& c : ext) c = (char)tolower(c);
        struct stat s;
        if ((ext == ".mtl" || ext == ".bin") && stat((dir + "/" + name).c_str(), &s) == 0 && S_ISREG(s.st_mode) && s.st_mtime > b.st_mtime) { fresh = false; break; }
    }
    closedir(d);
    return fresh;
}
bool textures_older_than_cache(const yrt_flat_scene& flat, const std::string& scene_path, const std::string& cache_path) {
    struct stat b, s;
    if (stat(cache_path.c_str(), &b) != 0) return false;
    const std::string dir = dir_of(scene_path);
    for (const auto& f : flat.tex_files)
        if (stat((dir + "/" + f).c_str(), &s) != 0 || s.st_mtime > b.st_mtime) return false;
    return true;
}
}  // namespace

int main(int argc, char** argv) {
    auto parser = yu::cmdline::make_parser(argc, argv, "raytrace", "raytrace scene");
    auto resolution = yu::cmdline::parse_opti(parser, "--resolution", "-r", "vertical resolution", 720);
    auto samples = yu::cmdline::parse_opti(parser, "--samples", "-s", "per-pixel samples", 1);
    auto amb = yu::cmdline::parse_optf(parser, "--ambient", "-a", "ambient color", 0.1f);
    auto imageout = yu::cmdline::parse_opts(parser, "--output", "-o", "output image", "out.png");
    // additive (not in the reference)
    auto gpus = yu::cmdline::parse_opti(parser, "--gpus", "-g", "number of GPUs (interleaved row tiles)", 1);
    auto verbose = yu::cmdline::parse_flag(parser, "--stats", "", "print ray counts and timings", false);
    auto allow_nonrigid = yu::cmdline::parse_flag(parser, "--allow-nonrigid", "", "ignored (scenes with scaled / sheared instance frames are always rendered)", false);
    auto use_cache = yu::cmdline::parse_flag(parser, "--cache", "", "reuse / write the flattened scene <scene>.yrts", false);
    auto device_ldr = yu::cmdline::parse_flag(parser, "--device-ldr", "", "tonemap on the GPU (8-bit outputs)", false);
    auto fast_png = yu::cmdline::parse_flag(parser, "--fast-png", "", "write .png outputs with the parallel encoder", false);
    auto scenein = yu::cmdline::parse_args(parser, "scenein", "input scene", "scene.obj", true);
    yu::cmdline::check_parser(parser);

    double t_start = now_ms();
    printf("loading scene %s\n", scenein.c_str());
    yrt_flat_scene flat;
    std::string err;
    std::string cache_path = scenein + ".yrts";
    bool from_cache = false;
    if (use_cache && cache_is_fresh(scenein, cache_path)) {
        from_cache = yrt_flat_load(cache_path, flat, err);
        if (!from_cache) printf("ignoring scene cache: %s\n", err.c_str());
        else if (!textures_older_than_cache(flat, scenein, cache_path)) from_cache = false;   // a texture changed since
    }
    double t_loaded = now_ms(), t_flat = t_loaded;
    if (!from_cache) {
        auto scn = load_scene(scenein);
        t_loaded = now_ms();
        if (!yrt_flatten(scn, flat, err)) {
            printf("could not flatten scene: %s\n", err.c_str());
            exit(1);
        }
        t_flat = now_ms();
        if (use_cache && !yrt_flat_save(flat, cache_path, err)) printf("could not write scene cache: %s\n", err.c_str());
    }
    if (!flat.has_camera) {
        printf("scene has no camera\n");
        exit(1);
    }

    printf("creating bvh\n");
    yrt_set_option("pin_host_frames", 1);   // the frame buffers below live until the process ends
    if (yrt_init(gpus) != YRT_OK) {
        printf("%s\n", yrt_last_error());
        exit(1);
    }
    yrt_scene_desc desc = flat.desc();
    (void)allow_nonrigid;   // scenes with scaled / sheared instance frames are traced through the reference's own instance tree (include/yrt_b200.h)
    if (verbose)
        if (int nonrigid = yrt_desc_nonrigid_instances(&desc)) printf("%d instance frame(s) are not rigid: tracing through the reference's instance tree\n", nonrigid);
    yrt_scene* gscn = nullptr;
    if (yrt_scene_create(&desc, &gscn) != YRT_OK) {
        printf("%s\n", yrt_last_error());
        exit(1);
    }
    int width = yrt_image_width(&flat.cam, resolution);   // (int)std::round(cam->aspect * resolution), src/raytrace.cpp:216
    bool want_hdr = imageout.length() >= 4 && imageout.substr(imageout.length() - 4) == ".hdr";   // src/image.cpp:82
    bool ldr_on_device = device_ldr && !want_hdr;
    // the frame buffers and the render workspace belong to the set-up, like the reference's image allocation (raytrace.cpp:217)
    auto hdr = ldr_on_device ? image4f() : image4f(width, resolution);
    auto ldr = ldr_on_device ? image4b(width, resolution) : image4b();
    if (yrt_scene_prepare(gscn, width, resolution, samples) != YRT_OK) {
        printf("%s\n", yrt_last_error());
        exit(1);
    }
    double t_built = now_ms();

    printf("tracing scene\n");
    float ambient[3] = {amb, amb, amb};
    yrt_stats st;
    memset(&st, 0, sizeof(st));
    yrt_stats* want = verbose ? &st : nullptr;     // (without per-kernel statistics the frame runs as two overlapping pipelines)
    int rc;
    if (ldr_on_device) {
        rc = yrt_render_ldr(gscn, &flat.cam, ambient, width, resolution, samples, (uint8_t*)ldr.pixels.data(), nullptr, want);
    } else {
        rc = yrt_render(gscn, &flat.cam, ambient, width, resolution, samples, (float*)hdr.pixels.data(), want);
    }
    if (rc != YRT_OK) {
        printf("%s\n", yrt_last_error());
        exit(1);
    }
    double t_traced = now_ms();
    // mirror bounces dropped at the recursion cap (only scenes with mirrors can have any; one small read per GPU)
    int64_t truncated = st.truncated_paths;
    bool mirrors = false;
    for (float kr : flat.mat_kr) mirrors = mirrors || kr > 0.0f;   // src/raytrace.cpp:190
    if (!verbose && mirrors && yrt_frame_truncated_paths(gscn, &truncated) != YRT_OK) truncated = 0;
    yrt_scene_destroy(gscn);

    printf("saving image %s\n", imageout.c_str());
    bool is_png = imageout.length() >= 4 && imageout.substr(imageout.length() - 4) == ".png";
    if (fast_png && is_png && !want_hdr) {
        if (!ldr_on_device) ldr = tonemap(hdr, 0, false);   // the reference's own host tonemap (src/image.cpp:55-78,:85)
        if (yrt_write_png(imageout.c_str(), (const uint8_t*)ldr.pixels.data(), ldr.width, ldr.height, 0, 0) != YRT_OK) {
            printf("%s\n", yrt_last_error());
            exit(1);
        }
    } else if (ldr_on_device) save_image(imageout, ldr);   // what save_hdr_or_ldr does after its host tonemap (src/image.cpp:85-86)
    else save_hdr_or_ldr(imageout, hdr);
    double t_saved = now_ms();

    if (truncated > 0)
        fprintf(stderr, "warning: %lld mirror bounce(s) were cut at the recursion cap of this path (YRT_MAX_DEPTH, default 64); the reference recurses without a bound\n",
                (long long)truncated);
    if (verbose) {
        long long rays = (long long)(st.primary_rays + st.reflection_rays + st.shadow_rays);
        double ms = t_traced - t_built;
        printf("rays %lld (primary %lld, reflection %lld, shadow %lld) on %d GPU(s): %.3f ms device, %.3f ms call, %.1f Mrays/s\n", rays,
               (long long)st.primary_rays, (long long)st.reflection_rays, (long long)st.shadow_rays, st.n_gpus, st.ms_total, ms,
               rays / (ms * 1e3));
        printf("phases ms: load %.1f%s, flatten %.1f, init+upload+bvh %.1f, trace %.1f, tonemap+save %.1f, total %.1f\n", t_loaded - t_start,
               from_cache ? " (scene cache)" : "", t_flat - t_loaded, t_built - t_flat, ms, t_saved - t_traced, t_saved - t_start);
    }
}
