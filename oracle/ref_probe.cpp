// ref_probe.cpp — TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference (compiled from
// /root/reference/src where it lies; nothing is copied) and dumps what parity tests compare against:
//   ref_probe ids    scene.obj R S out.bin   per primary ray: int32 instance index, int32 shape index,
//                                           int32 ei, float dist, float ew.y, float ew.z   (24 B / ray)
//   ref_probe image  scene.obj R S A out.bin float32 W*H*4 image4f of raytrace() BEFORE tonemap
//                                           (header: int32 W, int32 H), plus timing/ray counts on stdout
//   ref_probe brute  scene.obj R S out.bin   like `ids` but through intersect_scene (scene.cpp:362,
//                                           brute force: a BVH-independent oracle)
// The reference's raytrace.cpp is pulled in as a whole TU with its main() renamed, so eval_camera /
// shade / raytrace are the reference's own object code.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#define main reference_main
#include "raytrace.cpp"   // reference src/raytrace.cpp (found via -I)
#undef main

intersection3f intersect_scene(const scene* scn, const ray3f& ray);   // scene.cpp:362 (external linkage)

static int usage() {
    fprintf(stderr, "usage: ref_probe ids|brute scene.obj R S out.bin | ref_probe image scene.obj R S A out.bin\n");
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 6) return usage();
    std::string mode = argv[1];
    auto scn = load_scene(argv[2]);
    int resolution = atoi(argv[3]), samples = atoi(argv[4]);
    auto t0 = std::chrono::steady_clock::now();
    build_bvh(scn, false);
    auto t1 = std::chrono::steady_clock::now();
    auto cam = scn->cameras.front();
    int W = (int)std::round(cam->aspect * resolution), H = resolution;
    if (mode == "ids" || mode == "brute") {
        std::map<const instance*, int> iid;
        std::map<const shape*, int> sid;
        {   // explicit counters: `m[k] = m.size()` has unspecified evaluation order before C++17
            int n = 0;
            for (auto i : scn->instances) iid[i] = n++;
            n = 0;
            for (auto s : scn->shapes) sid[s] = n++;
        }
        FILE* f = fopen(argv[5], "wb");
        if (!f) return 1;
        int hdr[4] = {W, H, samples, 0};
        fwrite(hdr, 4, 4, f);
        // same sample positions as raytrace() (raytrace.cpp:228-241); written row-major, then (jj,ii)
        for (int j = 0; j < H; j++)
            for (int i = 0; i < W; i++)
                for (int jj = 0; jj < samples; jj++)
                    for (int ii = 0; ii < samples; ii++) {
                        vec2f uv = {(i + (ii + 0.5f) / samples) / W, (j + (jj + 0.5f) / samples) / H};
                        auto ray = eval_camera(cam, uv);
                        auto isec = mode == "ids" ? intersect_first(scn, ray) : intersect_scene(scn, ray);
                        struct { int inst, shape, ei; float dist, w1, w2; } rec;
                        if (isec.hit()) {
                            rec.inst = iid[isec.ist]; rec.shape = sid[isec.ist->shp]; rec.ei = isec.ei;
                            rec.dist = isec.dist; rec.w1 = isec.ew.y; rec.w2 = isec.ew.z;
                        } else {
                            rec.inst = rec.shape = rec.ei = -1; rec.dist = rec.w1 = rec.w2 = 0;
                        }
                        fwrite(&rec, sizeof(rec), 1, f);
                    }
        fclose(f);
        printf("{\"mode\": \"%s\", \"width\": %d, \"height\": %d, \"samples\": %d}\n", mode.c_str(), W, H, samples);
        return 0;
    }
    if (mode == "image") {
        if (argc < 7) return usage();
        float amb = (float)atof(argv[5]);
        auto t2 = std::chrono::steady_clock::now();
        auto img = raytrace(scn, vec3f{amb, amb, amb}, resolution, samples);
        auto t3 = std::chrono::steady_clock::now();
        FILE* f = fopen(argv[6], "wb");
        if (!f) return 1;
        int hdr[2] = {img.width, img.height};
        fwrite(hdr, 4, 2, f);
        fwrite(img.pixels.data(), sizeof(vec4f), img.pixels.size(), f);
        fclose(f);
        printf("{\"mode\": \"image\", \"width\": %d, \"height\": %d, \"samples\": %d, \"build_bvh_s\": %.6f, \"raytrace_s\": %.6f}\n",
               img.width, img.height, samples, std::chrono::duration<double>(t1 - t0).count(),
               std::chrono::duration<double>(t3 - t2).count());
        return 0;
    }
    return usage();
}
