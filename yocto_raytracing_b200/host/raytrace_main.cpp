// raytrace_main.cpp — the drop-in `raytrace` CLI: the reference's main() (src/raytrace.cpp:256-287)
// with build_bvh(scn,false) (:278) and raytrace(scn, amb, resolution, samples) (:282) replaced by the
// C ABI of include/yrt_b200.h.  Everything else is the reference's own code, linked unchanged:
// load_scene (src/scene.cpp:113), save_hdr_or_ldr (src/image.cpp:81), yu::cmdline
// (src/ext/yocto_utils.h:1085+).  Same flags, same four progress lines; additive flags only.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "ext/yocto_utils.h"   // reference
#include "scene.h"             // reference
#include "yrt_flatten.h"

int main(int argc, char** argv) {
    auto parser = yu::cmdline::make_parser(argc, argv, "raytrace", "raytrace scene");
    auto resolution = yu::cmdline::parse_opti(parser, "--resolution", "-r", "vertical resolution", 720);
    auto samples = yu::cmdline::parse_opti(parser, "--samples", "-s", "per-pixel samples", 1);
    auto amb = yu::cmdline::parse_optf(parser, "--ambient", "-a", "ambient color", 0.1f);
    auto imageout = yu::cmdline::parse_opts(parser, "--output", "-o", "output image", "out.png");
    // additive (not in the reference)
    auto gpus = yu::cmdline::parse_opti(parser, "--gpus", "-g", "number of GPUs (interleaved row tiles)", 1);
    auto verbose = yu::cmdline::parse_flag(parser, "--stats", "", "print ray counts and timings", false);
    auto scenein = yu::cmdline::parse_args(parser, "scenein", "input scene", "scene.obj", true);
    yu::cmdline::check_parser(parser);

    printf("loading scene %s\n", scenein.c_str());
    auto scn = load_scene(scenein);

    printf("creating bvh\n");
    yrt_flat_scene flat;
    std::string err;
    if (!yrt_flatten(scn, flat, err)) {
        printf("could not flatten scene: %s\n", err.c_str());
        exit(1);
    }
    if (!flat.has_camera) {
        printf("scene has no camera\n");
        exit(1);
    }
    if (yrt_init(gpus) != YRT_OK) {
        printf("%s\n", yrt_last_error());
        exit(1);
    }
    yrt_scene_desc desc = flat.desc();
    yrt_scene* gscn = nullptr;
    if (yrt_scene_create(&desc, &gscn) != YRT_OK) {
        printf("%s\n", yrt_last_error());
        exit(1);
    }

    printf("tracing scene\n");
    auto cam = scn->cameras.front();
    auto hdr = image4f((int)std::round(cam->aspect * resolution), resolution);   // src/raytrace.cpp:216
    float ambient[3] = {amb, amb, amb};
    yrt_stats st;
    auto t0 = std::chrono::steady_clock::now();
    if (yrt_render(gscn, &flat.cam, ambient, hdr.width, hdr.height, samples, (float*)hdr.pixels.data(), &st) != YRT_OK) {
        printf("%s\n", yrt_last_error());
        exit(1);
    }
    auto t1 = std::chrono::steady_clock::now();
    if (verbose) {
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        long long rays = (long long)(st.primary_rays + st.reflection_rays + st.shadow_rays);
        printf("rays %lld (primary %lld, reflection %lld, shadow %lld) on %d GPU(s): %.3f ms device, %.3f ms call, %.1f Mrays/s\n", rays,
               (long long)st.primary_rays, (long long)st.reflection_rays, (long long)st.shadow_rays, st.n_gpus, st.ms_total, ms,
               rays / (ms * 1e3));
    }
    yrt_scene_destroy(gscn);

    printf("saving image %s\n", imageout.c_str());
    save_hdr_or_ldr(imageout, hdr);
}
