// yrt_flatten.cpp — see yrt_flatten.h.  Compiled with -I <reference>/src (scene.h, vmath.h, image.h).
#include "yrt_flatten.h"

#include <cstdio>
#include <cstring>
#include <map>

#include "scene.h"   // the reference's src/scene.h

yrt_scene_desc yrt_flat_scene::desc() const {
    yrt_scene_desc d;
    memset(&d, 0, sizeof(d));
    d.n_shapes = (int32_t)shape_kind.size();
    d.n_instances = (int32_t)inst_shape.size();
    d.n_materials = (int32_t)mat_rs.size();
    d.n_textures = (int32_t)tex_w.size();
    d.n_verts = (int32_t)(pos.size() / 3);
    d.n_elem_idx = (int32_t)elem_idx.size();
    d.shape_kind = shape_kind.data();
    d.shape_elem_off = shape_elem_off.data();
    d.shape_elem_cnt = shape_elem_cnt.data();
    d.shape_vert_off = shape_vert_off.data();
    d.shape_vert_cnt = shape_vert_cnt.data();
    d.shape_has_uv = shape_has_uv.data();
    d.shape_has_radius = shape_has_radius.data();
    d.elem_idx = elem_idx.data();
    d.pos = pos.data();
    d.norm = norm.data();
    d.uv = uv.data();
    d.radius = radius.data();
    d.inst_frame = inst_frame.data();
    d.inst_shape = inst_shape.data();
    d.inst_mat = inst_mat.data();
    d.mat_ke = mat_ke.data();
    d.mat_kd = mat_kd.data();
    d.mat_ks = mat_ks.data();
    d.mat_kr = mat_kr.data();
    d.mat_rs = mat_rs.data();
    d.mat_kd_tex = mat_kd_tex.data();
    d.mat_ks_tex = mat_ks_tex.data();
    d.tex_w = tex_w.data();
    d.tex_h = tex_h.data();
    d.tex_off = tex_off.data();
    d.tex_rgba8 = tex_rgba8.data();
    d.tex_bytes = (int64_t)tex_rgba8.size();
    return d;
}

static void push3(std::vector<float>& v, const vec3f& a) { v.push_back(a.x); v.push_back(a.y); v.push_back(a.z); }

yrt_camera yrt_flatten_camera(const camera* cam) {
    yrt_camera c;
    const frame3f& f = cam->frame;
    const vec3f* ax[4] = {&f.x, &f.y, &f.z, &f.o};
    for (int k = 0; k < 4; k++) { c.frame[3 * k] = ax[k]->x; c.frame[3 * k + 1] = ax[k]->y; c.frame[3 * k + 2] = ax[k]->z; }
    c.fovy = cam->fovy; c.aspect = cam->aspect; c.aperture = cam->aperture; c.focus = cam->focus;
    return c;
}

bool yrt_flatten(const scene* scn, yrt_flat_scene& out, std::string& err) {
    out = yrt_flat_scene();
    std::map<const shape*, int> shape_id;
    std::map<const material*, int> mat_id;
    std::map<const texture*, int> tex_id;
    char msg[256];

    for (auto txt : scn->textures) {
        int id = (int)tex_id.size();
        tex_id[txt] = id;
        // eval_texture only ever reads texture::ldr (src/raytrace.cpp:43); an .hdr texture leaves it 0x0
        out.tex_files.push_back(txt->filename);
        out.tex_w.push_back(txt->ldr.width);
        out.tex_h.push_back(txt->ldr.height);
        out.tex_off.push_back((int64_t)out.tex_rgba8.size());
        const unsigned char* px = (const unsigned char*)txt->ldr.pixels.data();
        out.tex_rgba8.insert(out.tex_rgba8.end(), px, px + 4 * txt->ldr.pixels.size());
    }
    auto tex_of = [&](const texture* t) { return t ? tex_id.at(t) : -1; };

    for (auto mat : scn->materials) {
        mat_id[mat] = (int)mat_id.size();
        push3(out.mat_ke, mat->ke); push3(out.mat_kd, mat->kd); push3(out.mat_ks, mat->ks); push3(out.mat_kr, mat->kr);
        out.mat_rs.push_back(mat->rs);
        out.mat_kd_tex.push_back(tex_of(mat->kd_txt));
        out.mat_ks_tex.push_back(tex_of(mat->ks_txt));
    }

    for (auto shp : scn->shapes) {
        int sid = (int)shape_id.size();
        shape_id[shp] = sid;
        int kinds = (!shp->triangles.empty()) + (!shp->lines.empty()) + (!shp->points.empty());
        if (kinds > 1) {
            // intersect_bvh dispatches triangles > lines > points (scene.cpp:405-427) but eval_* dispatches
            // points > lines > triangles (scene.h:160-165): a mixed shape is inconsistent in the reference
            snprintf(msg, sizeof(msg), "shape %d ('%s') mixes element kinds", sid, shp->name.c_str());
            err = msg;
            return false;
        }
        int kind = !shp->triangles.empty() ? YRT_TRIANGLES : (!shp->lines.empty() ? YRT_LINES : YRT_POINTS);
        size_t nv = shp->pos.size();
        if (shp->norm.size() != nv) { snprintf(msg, sizeof(msg), "shape %d: %zu normals for %zu positions", sid, shp->norm.size(), nv); err = msg; return false; }
        bool has_uv = shp->texcoord.size() == nv && nv > 0;
        bool has_r = shp->radius.size() == nv && nv > 0;
        out.shape_kind.push_back(kind);
        out.shape_elem_off.push_back((int32_t)out.elem_idx.size());
        out.shape_vert_off.push_back((int32_t)(out.pos.size() / 3));
        out.shape_vert_cnt.push_back((int32_t)nv);
        out.shape_has_uv.push_back(has_uv);
        out.shape_has_radius.push_back(has_r);
        if (kind == YRT_TRIANGLES) {
            out.shape_elem_cnt.push_back((int32_t)shp->triangles.size());
            for (auto& t : shp->triangles) { out.elem_idx.push_back(t.x); out.elem_idx.push_back(t.y); out.elem_idx.push_back(t.z); }
        } else if (kind == YRT_LINES) {
            out.shape_elem_cnt.push_back((int32_t)shp->lines.size());
            for (auto& l : shp->lines) { out.elem_idx.push_back(l.x); out.elem_idx.push_back(l.y); }
        } else {
            out.shape_elem_cnt.push_back((int32_t)shp->points.size());
            for (auto p : shp->points) out.elem_idx.push_back(p);
        }
        for (size_t i = 0; i < nv; i++) {
            push3(out.pos, shp->pos[i]);
            push3(out.norm, shp->norm[i]);
            out.uv.push_back(has_uv ? shp->texcoord[i].x : 0.f);
            out.uv.push_back(has_uv ? shp->texcoord[i].y : 0.f);
            out.radius.push_back(has_r ? shp->radius[i] : 0.f);
        }
    }

    for (auto ist : scn->instances) {
        if (!shape_id.count(ist->shp) || !mat_id.count(ist->mat)) { err = "instance refers to a shape/material outside the scene"; return false; }
        const frame3f& f = ist->frame;
        push3(out.inst_frame, f.x); push3(out.inst_frame, f.y); push3(out.inst_frame, f.z); push3(out.inst_frame, f.o);
        out.inst_shape.push_back(shape_id[ist->shp]);
        out.inst_mat.push_back(mat_id[ist->mat]);
    }

    if (!scn->cameras.empty()) {
        out.cam = yrt_flatten_camera(scn->cameras.front());
        out.has_camera = true;
    }
    return true;
}

namespace {
struct Writer {
    FILE* f;
    int n = 0;
    void arr(const char* name, int dtype, const void* data, int64_t count, size_t elem) {
        char nm[24];
        memset(nm, 0, sizeof(nm));
        strncpy(nm, name, sizeof(nm) - 1);
        fwrite(nm, 1, sizeof(nm), f);
        int32_t dt = dtype;
        fwrite(&dt, 4, 1, f);
        fwrite(&count, 8, 1, f);
        size_t bytes = (size_t)count * elem;
        if (bytes) fwrite(data, 1, bytes, f);
        static const char pad[8] = {0};
        if (bytes % 8) fwrite(pad, 1, 8 - bytes % 8, f);
        n++;
    }
    void i32(const char* name, const std::vector<int32_t>& v) { arr(name, 0, v.data(), (int64_t)v.size(), 4); }
    void f32(const char* name, const std::vector<float>& v) { arr(name, 1, v.data(), (int64_t)v.size(), 4); }
};
}  // namespace

bool yrt_flat_save(const yrt_flat_scene& fs, const std::string& path, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot open " + path; return false; }
    fwrite("YRTSCN01", 1, 8, f);
    int32_t n_arrays = 26;
    fwrite(&n_arrays, 4, 1, f);
    int32_t zero = 0;
    fwrite(&zero, 4, 1, f);
    Writer w{f};
    w.i32("shape_kind", fs.shape_kind); w.i32("shape_elem_off", fs.shape_elem_off); w.i32("shape_elem_cnt", fs.shape_elem_cnt);
    w.i32("shape_vert_off", fs.shape_vert_off); w.i32("shape_vert_cnt", fs.shape_vert_cnt);
    w.i32("shape_has_uv", fs.shape_has_uv); w.i32("shape_has_radius", fs.shape_has_radius);
    w.i32("elem_idx", fs.elem_idx);
    w.f32("pos", fs.pos); w.f32("norm", fs.norm); w.f32("uv", fs.uv); w.f32("radius", fs.radius);
    w.f32("inst_frame", fs.inst_frame); w.i32("inst_shape", fs.inst_shape); w.i32("inst_mat", fs.inst_mat);
    w.f32("mat_ke", fs.mat_ke); w.f32("mat_kd", fs.mat_kd); w.f32("mat_ks", fs.mat_ks); w.f32("mat_kr", fs.mat_kr); w.f32("mat_rs", fs.mat_rs);
    w.i32("mat_kd_tex", fs.mat_kd_tex); w.i32("mat_ks_tex", fs.mat_ks_tex);
    w.i32("tex_w", fs.tex_w); w.i32("tex_h", fs.tex_h);
    w.arr("tex_off", 3, fs.tex_off.data(), (int64_t)fs.tex_off.size(), 8);
    w.arr("tex_rgba8", 2, fs.tex_rgba8.data(), (int64_t)fs.tex_rgba8.size(), 1);
    std::vector<float> cam(fs.cam.frame, fs.cam.frame + 12);
    cam.push_back(fs.cam.fovy); cam.push_back(fs.cam.aspect); cam.push_back(fs.cam.aperture); cam.push_back(fs.cam.focus);
    // camera rides as a 27th array; the count in the header is patched below
    w.f32("camera", cam);
    {   // file names of the textures, one per line: what the cache depends on besides the scene file itself
        std::string names;
        for (const auto& n : fs.tex_files) names += n + "\n";
        w.arr("tex_files", 2, names.data(), (int64_t)names.size(), 1);
    }
    fseek(f, 8, SEEK_SET);
    int32_t n = w.n;
    fwrite(&n, 4, 1, f);
    bool ok = !ferror(f);
    fclose(f);
    if (!ok) err = "write error on " + path;
    return ok;
}

bool yrt_flat_load(const std::string& path, yrt_flat_scene& fs, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    char magic[8];
    int32_t n_arrays = 0, reserved = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "YRTSCN01", 8) == 0 && fread(&n_arrays, 4, 1, f) == 1 && fread(&reserved, 4, 1, f) == 1;
    if (!ok) { fclose(f); err = path + ": not a YRTSCN01 file"; return false; }
    fs = yrt_flat_scene();
    std::vector<float> cam;
    std::string tex_names;
    // no count may promise more bytes than the file still holds (a corrupt cache must not cause a huge allocation)
    long file_size = 0;
    {
        long here = ftell(f);
        fseek(f, 0, SEEK_END);
        file_size = ftell(f);
        fseek(f, here, SEEK_SET);
    }
    ok = n_arrays >= 0 && n_arrays <= 1024;
    for (int a = 0; a < n_arrays && ok; a++) {
        char nm[25] = {0};
        int32_t dt = 0;
        int64_t count = 0;
        ok = fread(nm, 1, 24, f) == 24 && fread(&dt, 4, 1, f) == 1 && fread(&count, 8, 1, f) == 1 && count >= 0 && dt >= 0 && dt <= 3;
        if (!ok) break;
        static const size_t esz[4] = {4, 4, 1, 8};
        if ((unsigned long long)count > (unsigned long long)(file_size - ftell(f)) / esz[dt]) { ok = false; break; }
        size_t bytes = (size_t)count * esz[dt];
        std::string name(nm);
        void* dst = nullptr;
        auto want = [&](const char* n, int t) { return name == n && dt == t; };
#define YRT_I32(field) if (want(#field, 0)) { fs.field.resize((size_t)count); dst = fs.field.data(); }
#define YRT_F32(field) if (want(#field, 1)) { fs.field.resize((size_t)count); dst = fs.field.data(); }
        YRT_I32(shape_kind) YRT_I32(shape_elem_off) YRT_I32(shape_elem_cnt) YRT_I32(shape_vert_off) YRT_I32(shape_vert_cnt)
        YRT_I32(shape_has_uv) YRT_I32(shape_has_radius) YRT_I32(elem_idx) YRT_I32(inst_shape) YRT_I32(inst_mat)
        YRT_I32(mat_kd_tex) YRT_I32(mat_ks_tex) YRT_I32(tex_w) YRT_I32(tex_h)
        YRT_F32(pos) YRT_F32(norm) YRT_F32(uv) YRT_F32(radius) YRT_F32(inst_frame)
        YRT_F32(mat_ke) YRT_F32(mat_kd) YRT_F32(mat_ks) YRT_F32(mat_kr) YRT_F32(mat_rs)
#undef YRT_I32
#undef YRT_F32
        if (want("tex_off", 3)) { fs.tex_off.resize((size_t)count); dst = fs.tex_off.data(); }
        if (want("tex_rgba8", 2)) { fs.tex_rgba8.resize((size_t)count); dst = fs.tex_rgba8.data(); }
        if (want("camera", 1)) { cam.resize((size_t)count); dst = cam.data(); }
        if (want("tex_files", 2)) { tex_names.resize((size_t)count); dst = &tex_names[0]; }
        if (dst) ok = bytes == 0 || fread(dst, 1, bytes, f) == bytes;
        else ok = fseek(f, (long)bytes, SEEK_CUR) == 0;   // unknown array: skip
        if (ok && bytes % 8) ok = fseek(f, (long)(8 - bytes % 8), SEEK_CUR) == 0;
    }
    fclose(f);
    if (!ok) { err = path + ": truncated or malformed"; return false; }
    for (size_t a0 = 0; a0 < tex_names.size();) {
        size_t e = tex_names.find('\n', a0);
        if (e == std::string::npos) e = tex_names.size();
        fs.tex_files.push_back(tex_names.substr(a0, e - a0));
        a0 = e + 1;
    }
    if (cam.size() == 16) {
        memcpy(fs.cam.frame, cam.data(), 12 * sizeof(float));
        fs.cam.fovy = cam[12]; fs.cam.aspect = cam[13]; fs.cam.aperture = cam[14]; fs.cam.focus = cam[15];
        fs.has_camera = true;
    }
    return true;
}
