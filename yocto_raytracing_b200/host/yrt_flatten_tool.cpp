// yrt_flatten_tool.cpp — `yrt_flatten scene.obj out.yrts`: load an OBJ with the reference's own loader
// (load_scene, src/scene.cpp:113) and write the flattened scene container the Python side reads.
// `yrt_flatten --check file.yrts`: read a container back with the CLI's loader (yrt_flat_load) and print its counts.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scene.h"   // reference
#include "yrt_flatten.h"

int main(int argc, char** argv) {
    if (argc == 3 && !strcmp(argv[1], "--check")) {
        yrt_flat_scene flat;
        std::string err;
        if (!yrt_flat_load(argv[2], flat, err)) {
            fprintf(stderr, "%s\n", err.c_str());
            return 1;
        }
        printf("%s: %zu shapes, %zu instances, %zu materials, %zu textures, %zu vertices, %zu index ints\n", argv[2], flat.shape_kind.size(),
               flat.inst_shape.size(), flat.mat_rs.size(), flat.tex_w.size(), flat.pos.size() / 3, flat.elem_idx.size());
        return 0;
    }
    if (argc != 3) {
        fprintf(stderr, "usage: %s scene.obj out.yrts\n", argv[0]);
        return 2;
    }
    auto scn = load_scene(argv[1]);
    yrt_flat_scene flat;
    std::string err;
    if (!yrt_flatten(scn, flat, err)) {
        fprintf(stderr, "could not flatten scene: %s\n", err.c_str());
        return 1;
    }
    if (!yrt_flat_save(flat, argv[2], err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    printf("%s: %zu shapes, %zu instances, %zu materials, %zu textures, %zu vertices, %zu index ints\n", argv[2], flat.shape_kind.size(),
           flat.inst_shape.size(), flat.mat_rs.size(), flat.tex_w.size(), flat.pos.size() / 3, flat.elem_idx.size());
    return 0;
}
