// yrt_flatten.h — host shim between the reference's loaded `scene` (src/scene.h:136-155, built by
// load_scene, src/scene.cpp:113) and the SoA description the C ABI takes (include/yrt_b200.h).
//
// This is the only code that touches the reference's pointer graph; it is compiled against the
// reference's own headers (-I <reference>/src) and is what a maintainer adds next to main().
#ifndef YRT_FLATTEN_H_
#define YRT_FLATTEN_H_

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/yrt_b200.h"

struct scene;    // reference type, src/scene.h:136
struct camera;   // reference type, src/scene.h:115

struct yrt_flat_scene {
    std::vector<int32_t> shape_kind, shape_elem_off, shape_elem_cnt, shape_vert_off, shape_vert_cnt, shape_has_uv, shape_has_radius;
    std::vector<int32_t> elem_idx;
    std::vector<float> pos, norm, uv, radius;
    std::vector<float> inst_frame;
    std::vector<int32_t> inst_shape, inst_mat;
    std::vector<float> mat_ke, mat_kd, mat_ks, mat_kr, mat_rs;
    std::vector<int32_t> mat_kd_tex, mat_ks_tex;
    std::vector<int32_t> tex_w, tex_h;
    std::vector<int64_t> tex_off;
    std::vector<uint8_t> tex_rgba8;
    std::vector<std::string> tex_files;   // texture::filename of every texture (relative to the scene's directory): the cache's dependencies
    yrt_camera cam;          // scn->cameras.front() (src/raytrace.cpp:215)
    bool has_camera = false;

    // view over the vectors above; valid while this object is alive and unmodified
    yrt_scene_desc desc() const;
};

// flatten a loaded reference scene; false + message on scenes the path does not cover
bool yrt_flatten(const scene* scn, yrt_flat_scene& out, std::string& err);
yrt_camera yrt_flatten_camera(const camera* cam);

// "YRTSCN01" container (named typed arrays) read by yocto_raytracing_b200/scene.py
bool yrt_flat_save(const yrt_flat_scene& fs, const std::string& path, std::string& err);
// ... and read back: the CLI's scene cache (parsing the 3.3 MB instance10000 OBJ costs more than a frame)
bool yrt_flat_load(const std::string& path, yrt_flat_scene& fs, std::string& err);

#endif
