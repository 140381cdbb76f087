/* rr_mesh_io.cpp — triangle-soup ingest for rr_set_mesh: the step BEFORE the hot path.
 *
 * Replaces, for the formats the reference's launch files name, `rm::import_embree_map(map_file)`
 * (src/radar_simulator.cpp:149,164; launch/mulran_sim.launch:7 loads a single-mesh .ply, config/oru4.yaml:46-65 lists the
 * 18 scene-graph objects of a .dae). Rmagine delegates to assimp; neither is in this image, so the readers here are
 * written from the public format descriptions:
 *   .ply  ascii / binary_little_endian / binary_big_endian; `vertex` x y z (any scalar type, extra properties skipped),
 *         `face` with a vertex_indices|vertex_index list (polygons are fan-triangulated); other elements skipped.
 *         One mesh -> object id 0 for every face (a single-mesh .ply gives obj_id = 0).
 *   .obj  v / f (v, v/vt, v/vt/vn, v//vn, negative indices); every `o` or `g` statement opens the next object id, which is
 *         how a scene graph exported to .obj keeps the per-object ids that index `object_materials` (RadarCPU.cpp:268).
 * Host code only: nothing here touches the device.
 */
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/radarays_b200.h"

namespace {

struct Soup {
    std::vector<float> v;
    std::vector<uint32_t> t, o;
    uint32_t n_objects = 0;
};

bool ends_with_ci(const std::string& s, const char* suf)
{
    const size_t n = strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; i++)
        if (tolower((unsigned char)s[s.size() - n + i]) != tolower((unsigned char)suf[i])) return false;
    return true;
}

/* ---------------------------------------------------------------- PLY */
enum PlyType { T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_BAD };

PlyType ply_type(const std::string& s)
{
    if (s == "char" || s == "int8") return T_I8;
    if (s == "uchar" || s == "uint8") return T_U8;
    if (s == "short" || s == "int16") return T_I16;
    if (s == "ushort" || s == "uint16") return T_U16;
    if (s == "int" || s == "int32") return T_I32;
    if (s == "uint" || s == "uint32") return T_U32;
    if (s == "float" || s == "float32") return T_F32;
    if (s == "double" || s == "float64") return T_F64;
    return T_BAD;
}
size_t ply_size(PlyType t) { static const size_t sz[] = {1, 1, 2, 2, 4, 4, 4, 8, 0}; return sz[t]; }

struct PlyProp { std::string name; bool is_list = false; PlyType count_type = T_BAD, type = T_BAD; };
struct PlyElem { std::string name; size_t count = 0; std::vector<PlyProp> props; };

double ply_read_bin(const unsigned char*& p, const unsigned char* end, PlyType t, bool swap, bool& ok)
{
    const size_t n = ply_size(t);
    if ((size_t)(end - p) < n) { ok = false; return 0.0; }
    unsigned char b[8];
    for (size_t i = 0; i < n; i++) b[i] = swap ? p[n - 1 - i] : p[i];
    p += n;
    switch (t) {
        case T_I8: { int8_t x; memcpy(&x, b, 1); return x; }
        case T_U8: { uint8_t x; memcpy(&x, b, 1); return x; }
        case T_I16: { int16_t x; memcpy(&x, b, 2); return x; }
        case T_U16: { uint16_t x; memcpy(&x, b, 2); return x; }
        case T_I32: { int32_t x; memcpy(&x, b, 4); return x; }
        case T_U32: { uint32_t x; memcpy(&x, b, 4); return x; }
        case T_F32: { float x; memcpy(&x, b, 4); return x; }
        case T_F64: { double x; memcpy(&x, b, 8); return x; }
        default: ok = false; return 0.0;
    }
}

void fan(Soup& s, const std::vector<long long>& idx, uint32_t obj, size_t n_verts, bool& ok)
{
    for (long long i : idx) if (i < 0 || (size_t)i >= n_verts) { ok = false; return; }
    for (size_t k = 1; k + 1 < idx.size(); k++) {
        s.t.push_back((uint32_t)idx[0]); s.t.push_back((uint32_t)idx[k]); s.t.push_back((uint32_t)idx[k + 1]);
        s.o.push_back(obj);
    }
}

bool load_ply(const std::string& path, Soup& s, std::string& err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    auto next_line = [&](std::string& line) -> bool {
        if (pos >= data.size()) return false;
        size_t e = data.find('\n', pos);
        if (e == std::string::npos) e = data.size();
        line = data.substr(pos, e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e + 1;
        return true;
    };
    std::string line;
    if (!next_line(line) || line != "ply") { err = "not a PLY file (magic)"; return false; }
    int fmt = -1;                                  /* 0 ascii, 1 little, 2 big */
    std::vector<PlyElem> elems;
    bool header_done = false;
    while (next_line(line)) {
        std::istringstream is(line);
        std::string kw; is >> kw;
        if (kw == "format") {
            std::string v; is >> v;
            fmt = (v == "ascii") ? 0 : (v == "binary_little_endian") ? 1 : (v == "binary_big_endian") ? 2 : -1;
        } else if (kw == "element") {
            PlyElem e; is >> e.name >> e.count; elems.push_back(e);
        } else if (kw == "property") {
            if (elems.empty()) { err = "PLY: property before element"; return false; }
            PlyProp p; std::string a; is >> a;
            if (a == "list") { std::string c, t; is >> c >> t >> p.name; p.is_list = true; p.count_type = ply_type(c); p.type = ply_type(t); }
            else { p.type = ply_type(a); is >> p.name; }
            if (p.type == T_BAD || (p.is_list && p.count_type == T_BAD)) { err = "PLY: unknown property type in '" + line + "'"; return false; }
            elems.back().props.push_back(p);
        } else if (kw == "end_header") { header_done = true; break; }
    }
    if (!header_done || fmt < 0) { err = "PLY: bad header (format / end_header)"; return false; }
    const uint16_t probe = 1; const bool host_little = *reinterpret_cast<const unsigned char*>(&probe) == 1;
    const bool swap = (fmt == 1 && !host_little) || (fmt == 2 && host_little);
    const unsigned char* bp = reinterpret_cast<const unsigned char*>(data.data()) + pos;
    const unsigned char* bend = reinterpret_cast<const unsigned char*>(data.data()) + data.size();
    std::istringstream as;
    if (fmt == 0) as.str(data.substr(pos));
    bool ok = true;
    auto scalar = [&](PlyType t) -> double {
        if (fmt == 0) { double x = 0; if (!(as >> x)) ok = false; return x; }
        return ply_read_bin(bp, bend, t, swap, ok);
    };
    size_t n_verts = 0;
    bool have_vertex = false;
    std::vector<long long> idx;
    for (const PlyElem& e : elems) {
        int ix = -1, iy = -1, iz = -1, il = -1;
        for (size_t k = 0; k < e.props.size(); k++) {
            const PlyProp& p = e.props[k];
            if (!p.is_list && p.name == "x") ix = (int)k;
            if (!p.is_list && p.name == "y") iy = (int)k;
            if (!p.is_list && p.name == "z") iz = (int)k;
            if (p.is_list && (p.name == "vertex_indices" || p.name == "vertex_index")) il = (int)k;
        }
        const bool is_vertex = (e.name == "vertex"), is_face = (e.name == "face");
        if (is_vertex && (ix < 0 || iy < 0 || iz < 0)) { err = "PLY: vertex element without x/y/z"; return false; }
        if (is_face && il < 0) { err = "PLY: face element without vertex_indices"; return false; }
        if (is_face && !have_vertex) { err = "PLY: face element before vertex element"; return false; }
        if (is_vertex) { s.v.resize(e.count * 3); n_verts = e.count; have_vertex = true; }
        for (size_t r = 0; r < e.count && ok; r++) {
            for (size_t k = 0; k < e.props.size() && ok; k++) {
                const PlyProp& p = e.props[k];
                if (!p.is_list) {
                    const double x = scalar(p.type);
                    if (is_vertex) {
                        if ((int)k == ix) s.v[3 * r] = (float)x;
                        else if ((int)k == iy) s.v[3 * r + 1] = (float)x;
                        else if ((int)k == iz) s.v[3 * r + 2] = (float)x;
                    }
                } else {
                    const double cnt = scalar(p.count_type);
                    if (!ok || cnt < 0 || cnt > 1e6) { ok = false; break; }
                    idx.resize((size_t)cnt);
                    for (size_t q = 0; q < idx.size(); q++) idx[q] = (long long)scalar(p.type);
                    if (ok && is_face && (int)k == il) fan(s, idx, 0u, n_verts, ok);
                }
            }
        }
        if (!ok) { err = "PLY: truncated or malformed '" + e.name + "' element (or a face index outside the vertex array)"; return false; }
    }
    if (!have_vertex) { err = "PLY: no vertex element"; return false; }
    s.n_objects = 1;
    return true;
}

/* ---------------------------------------------------------------- OBJ */
bool load_obj(const std::string& path, Soup& s, std::string& err)
{
    std::ifstream f(path);
    if (!f) { err = "cannot open " + path; return false; }
    std::string line;
    uint32_t obj = 0; bool obj_used = false, any_group = false;
    std::vector<long long> idx;
    size_t lineno = 0;
    while (std::getline(f, line)) {
        lineno++;
        std::istringstream is(line);
        std::string kw; is >> kw;
        if (kw == "v") {
            float x, y, z;
            if (!(is >> x >> y >> z)) { err = "OBJ: bad vertex at line " + std::to_string(lineno); return false; }
            s.v.push_back(x); s.v.push_back(y); s.v.push_back(z);
        } else if (kw == "o" || kw == "g") {
            if (obj_used || any_group) { if (obj_used) obj++; }
            any_group = true; obj_used = false;
        } else if (kw == "f") {
            idx.clear();
            std::string tok;
            const long long nv = (long long)(s.v.size() / 3);
            while (is >> tok) {
                const long long i = strtoll(tok.c_str(), nullptr, 10);      /* stops at '/' */
                if (i == 0) { err = "OBJ: bad face index at line " + std::to_string(lineno); return false; }
                idx.push_back(i > 0 ? i - 1 : nv + i);
            }
            bool ok = true;
            fan(s, idx, obj, (size_t)nv, ok);
            if (!ok) { err = "OBJ: face index outside the vertex array at line " + std::to_string(lineno); return false; }
            obj_used = true;
        }
    }
    s.n_objects = obj + 1;
    return true;
}

void set_err(char* err, size_t cap, const std::string& msg)
{
    if (err && cap) { snprintf(err, cap, "%s", msg.c_str()); }
}

}  // namespace

extern "C" {

int rr_mesh_load(const char* path, rr_mesh* out, char* err, size_t err_cap)
{
    if (!path || !out) { set_err(err, err_cap, "rr_mesh_load: NULL argument"); return RR_ERR_INVALID_ARGUMENT; }
    memset(out, 0, sizeof(*out));
    Soup s; std::string e;
    const std::string p(path);
    bool ok;
    if (ends_with_ci(p, ".ply")) ok = load_ply(p, s, e);
    else if (ends_with_ci(p, ".obj")) ok = load_obj(p, s, e);
    else { set_err(err, err_cap, "rr_mesh_load: unsupported mesh format (supported: .ply, .obj): " + p); return RR_ERR_INVALID_ARGUMENT; }
    if (!ok) { set_err(err, err_cap, e); return RR_ERR_INVALID_ARGUMENT; }
    if (s.t.empty()) { set_err(err, err_cap, "mesh file holds no faces: " + p); return RR_ERR_INVALID_ARGUMENT; }
    out->n_verts = s.v.size() / 3; out->n_tris = s.t.size() / 3; out->n_objects = s.n_objects;
    out->verts_xyz = (float*)malloc(s.v.size() * sizeof(float));
    out->tri_idx = (uint32_t*)malloc(s.t.size() * sizeof(uint32_t));
    out->tri_object_id = (uint32_t*)malloc(s.o.size() * sizeof(uint32_t));
    if (!out->verts_xyz || !out->tri_idx || !out->tri_object_id) {
        rr_mesh_free(out);
        set_err(err, err_cap, "rr_mesh_load: out of memory");
        return RR_ERR_INVALID_ARGUMENT;
    }
    memcpy(out->verts_xyz, s.v.data(), s.v.size() * sizeof(float));
    memcpy(out->tri_idx, s.t.data(), s.t.size() * sizeof(uint32_t));
    memcpy(out->tri_object_id, s.o.data(), s.o.size() * sizeof(uint32_t));
    return RR_OK;
}

void rr_mesh_free(rr_mesh* m)
{
    if (!m) return;
    free(m->verts_xyz); free(m->tri_idx); free(m->tri_object_id);
    memset(m, 0, sizeof(*m));
}

}  // extern "C"
