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
 *   .dae  COLLADA 1.4 as Blender writes it (the ORU map, launch/mro_husky.launch:4): library_geometries (POSITION source,
 *         triangles / polylist / polygons, fan-triangulated), the instantiated visual scene with nested nodes
 *         (matrix / translate / rotate / scale, composed in document order), instance_geometry, instance_node. Every
 *         mesh instance becomes the next object id in depth-first scene-graph order — the ids config/oru4.yaml:46-65
 *         lists; vertices are baked into the scene frame in fp64 and rounded once. Materials, normals, UVs, cameras,
 *         lights and <asset> up_axis / unit are ignored (the file's own axes are kept).
 * Host code only: nothing here touches the device.
 */
#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <new>
#include <stdexcept>
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
        /* an element can not hold more records than the rest of the file has bytes (ascii: >= 2 bytes per record,
         * binary: >= 1): bound the count BEFORE sizing anything by it */
        const size_t bytes_left = (fmt == 0) ? data.size() - pos : (size_t)(bend - bp);
        if (e.count > bytes_left) { err = "PLY: element '" + e.name + "' declares more records than the file holds"; return false; }
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

/* ---------------------------------------------------------------- COLLADA (.dae) */
struct XmlNode {
    std::string name, text;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<XmlNode> kids;
    const std::string* attr(const char* k) const { for (auto& a : attrs) if (a.first == k) return &a.second; return nullptr; }
    const XmlNode* child(const char* n) const { for (auto& c : kids) if (c.name == n) return &c; return nullptr; }
};

/* minimal non-validating XML reader: elements, attributes, character data; skips <?..?>, <!--..-->, <!..>, CDATA kept raw */
struct XmlReader {
    const std::string& d; size_t p = 0; std::string err;
    explicit XmlReader(const std::string& data) : d(data) {}
    void skip_ws() { while (p < d.size() && isspace((unsigned char)d[p])) p++; }
    bool skip_misc()
    {
        for (;;) {
            skip_ws();
            if (d.compare(p, 4, "<!--") == 0) { size_t e = d.find("-->", p + 4); if (e == std::string::npos) return false; p = e + 3; }
            else if (d.compare(p, 2, "<?") == 0) { size_t e = d.find("?>", p + 2); if (e == std::string::npos) return false; p = e + 2; }
            else if (d.compare(p, 2, "<!") == 0 && d.compare(p, 9, "<![CDATA[") != 0) { size_t e = d.find('>', p); if (e == std::string::npos) return false; p = e + 1; }
            else return true;
        }
    }
    bool parse(XmlNode& n)
    {
        if (!skip_misc() || p >= d.size() || d[p] != '<') { err = "XML: element expected"; return false; }
        p++;
        size_t s0 = p;
        while (p < d.size() && !isspace((unsigned char)d[p]) && d[p] != '>' && d[p] != '/') p++;
        n.name = d.substr(s0, p - s0);
        for (;;) {
            skip_ws();
            if (p >= d.size()) { err = "XML: unterminated tag"; return false; }
            if (d[p] == '/') { if (p + 1 < d.size() && d[p + 1] == '>') { p += 2; return true; } err = "XML: bad tag"; return false; }
            if (d[p] == '>') { p++; break; }
            size_t k0 = p;
            while (p < d.size() && d[p] != '=' && !isspace((unsigned char)d[p])) p++;
            std::string key = d.substr(k0, p - k0);
            skip_ws();
            if (p >= d.size() || d[p] != '=') { err = "XML: attribute without value"; return false; }
            p++; skip_ws();
            if (p >= d.size() || (d[p] != '"' && d[p] != '\'')) { err = "XML: unquoted attribute"; return false; }
            const char q = d[p++];
            size_t v0 = p;
            while (p < d.size() && d[p] != q) p++;
            if (p >= d.size()) { err = "XML: unterminated attribute"; return false; }
            n.attrs.emplace_back(key, d.substr(v0, p - v0));
            p++;
        }
        for (;;) {                                     /* content */
            size_t t0 = p;
            while (p < d.size() && d[p] != '<') p++;
            n.text.append(d, t0, p - t0);
            if (p >= d.size()) { err = "XML: missing </" + n.name + ">"; return false; }
            if (d.compare(p, 2, "</") == 0) {
                size_t e = d.find('>', p);
                if (e == std::string::npos) { err = "XML: bad end tag"; return false; }
                p = e + 1;
                return true;
            }
            if (d.compare(p, 4, "<!--") == 0 || d.compare(p, 2, "<?") == 0) { if (!skip_misc()) { err = "XML: bad comment"; return false; } continue; }
            if (d.compare(p, 9, "<![CDATA[") == 0) {
                size_t e = d.find("]]>", p + 9);
                if (e == std::string::npos) { err = "XML: bad CDATA"; return false; }
                n.text.append(d, p + 9, e - p - 9); p = e + 3; continue;
            }
            n.kids.emplace_back();
            if (!parse(n.kids.back())) return false;
        }
    }
};

struct Mat4 { double m[16]; };                          /* row-major, column vectors: p' = M p (COLLADA <matrix>) */
Mat4 mat_identity() { Mat4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0; return r; }
Mat4 mat_mul(const Mat4& a, const Mat4& b)
{
    Mat4 r{};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += a.m[4 * i + k] * b.m[4 * k + j]; r.m[4 * i + j] = s; }
    return r;
}
void parse_doubles(const std::string& t, std::vector<double>& out)
{
    const char* c = t.c_str(); char* e = nullptr;
    for (;;) { const double v = strtod(c, &e); if (e == c) break; out.push_back(v); c = e; }
}
void parse_ints(const std::string& t, std::vector<long long>& out)
{
    const char* c = t.c_str(); char* e = nullptr;
    for (;;) { const long long v = strtoll(c, &e, 10); if (e == c) break; out.push_back(v); c = e; }
}
std::string strip_hash(const std::string& s) { return (!s.empty() && s[0] == '#') ? s.substr(1) : s; }

struct DaeGeom { std::vector<float> v; std::vector<uint32_t> t; };     /* local-space triangle soup of one <geometry> */

bool dae_geometry(const XmlNode& g, DaeGeom& out, std::string& err)
{
    const XmlNode* mesh = g.child("mesh");
    if (!mesh) return true;                              /* splines, convex_mesh: no triangles */
    std::vector<std::pair<std::string, std::vector<double>>> sources;   /* id -> floats (xyz triples assumed for POSITION) */
    for (const XmlNode& s : mesh->kids) {
        if (s.name != "source") continue;
        const XmlNode* fa = s.child("float_array");
        const std::string* id = s.attr("id");
        if (!fa || !id) continue;
        sources.emplace_back(*id, std::vector<double>());
        parse_doubles(fa->text, sources.back().second);
    }
    std::string pos_src;
    std::string verts_id;
    if (const XmlNode* vs = mesh->child("vertices")) {
        if (const std::string* id = vs->attr("id")) verts_id = *id;
        for (const XmlNode& in : vs->kids)
            if (in.name == "input") { const std::string* sem = in.attr("semantic"); const std::string* src = in.attr("source"); if (sem && src && *sem == "POSITION") pos_src = strip_hash(*src); }
    }
    const std::vector<double>* pos = nullptr;
    for (auto& s : sources) if (s.first == pos_src) pos = &s.second;
    if (!pos || pos->size() % 3 != 0) { err = "COLLADA: geometry without a POSITION source of xyz triples"; return false; }
    out.v.resize(pos->size());
    for (size_t i = 0; i < pos->size(); i++) out.v[i] = (float)(*pos)[i];
    const size_t nv = pos->size() / 3;
    for (const XmlNode& prim : mesh->kids) {
        const bool tri = prim.name == "triangles", plist = prim.name == "polylist", polys = prim.name == "polygons";
        if (!tri && !plist && !polys) continue;
        int v_off = -1, stride = 0;
        for (const XmlNode& in : prim.kids) {
            if (in.name != "input") continue;
            const std::string* off = in.attr("offset"); const std::string* sem = in.attr("semantic");
            const int o = off ? atoi(off->c_str()) : 0;
            stride = std::max(stride, o + 1);
            if (sem && *sem == "VERTEX") v_off = o;
        }
        if (v_off < 0 || stride < 1) { err = "COLLADA: primitive without a VERTEX input"; return false; }
        std::vector<long long> p, vcount, idx;
        auto emit = [&](const std::vector<long long>& poly) -> bool {
            for (long long i : poly) if (i < 0 || (size_t)i >= nv) { err = "COLLADA: vertex index outside the position array"; return false; }
            for (size_t k = 1; k + 1 < poly.size(); k++) { out.t.push_back((uint32_t)poly[0]); out.t.push_back((uint32_t)poly[k]); out.t.push_back((uint32_t)poly[k + 1]); }
            return true;
        };
        if (polys) {
            for (const XmlNode& pe : prim.kids) {
                if (pe.name != "p") continue;
                p.clear(); parse_ints(pe.text, p);
                idx.clear();
                for (size_t k = 0; k + stride <= p.size(); k += stride) idx.push_back(p[k + v_off]);
                if (!emit(idx)) return false;
            }
            continue;
        }
        const XmlNode* pe = prim.child("p");
        if (!pe) continue;
        parse_ints(pe->text, p);
        if (plist) { if (const XmlNode* vc = prim.child("vcount")) parse_ints(vc->text, vcount); }
        size_t k = 0;
        if (tri) {
            for (; k + 3 * (size_t)stride <= p.size(); k += 3 * stride) {
                idx = {p[k + v_off], p[k + stride + v_off], p[k + 2 * stride + v_off]};
                if (!emit(idx)) return false;
            }
        } else {
            for (long long n : vcount) {
                if (n < 0 || k + (size_t)n * stride > p.size()) { err = "COLLADA: polylist shorter than its vcount"; return false; }
                idx.clear();
                for (long long q = 0; q < n; q++) idx.push_back(p[k + (size_t)q * stride + v_off]);
                k += (size_t)n * stride;
                if (!emit(idx)) return false;
            }
        }
    }
    return true;
}

Mat4 dae_node_transform(const XmlNode& node)
{
    Mat4 M = mat_identity();
    for (const XmlNode& t : node.kids) {                /* transforms compose in document order */
        std::vector<double> f;
        if (t.name == "matrix") {
            parse_doubles(t.text, f);
            if (f.size() == 16) { Mat4 A; for (int i = 0; i < 16; i++) A.m[i] = f[i]; M = mat_mul(M, A); }
        } else if (t.name == "translate") {
            parse_doubles(t.text, f);
            if (f.size() == 3) { Mat4 A = mat_identity(); A.m[3] = f[0]; A.m[7] = f[1]; A.m[11] = f[2]; M = mat_mul(M, A); }
        } else if (t.name == "scale") {
            parse_doubles(t.text, f);
            if (f.size() == 3) { Mat4 A = mat_identity(); A.m[0] = f[0]; A.m[5] = f[1]; A.m[10] = f[2]; M = mat_mul(M, A); }
        } else if (t.name == "rotate") {
            parse_doubles(t.text, f);
            if (f.size() == 4) {
                double x = f[0], y = f[1], z = f[2]; const double n = std::sqrt(x * x + y * y + z * z);
                if (n > 0) {
                    x /= n; y /= n; z /= n;
                    const double a = f[3] * M_PI / 180.0, c = std::cos(a), s = std::sin(a), C = 1 - c;
                    Mat4 A = mat_identity();
                    A.m[0] = c + x * x * C;     A.m[1] = x * y * C - z * s; A.m[2] = x * z * C + y * s;
                    A.m[4] = y * x * C + z * s; A.m[5] = c + y * y * C;     A.m[6] = y * z * C - x * s;
                    A.m[8] = z * x * C - y * s; A.m[9] = z * y * C + x * s; A.m[10] = c + z * z * C;
                    M = mat_mul(M, A);
                }
            }
        }
    }
    return M;
}

struct DaeDoc {
    std::vector<std::pair<std::string, DaeGeom>> geoms;
    std::vector<std::pair<std::string, const XmlNode*>> lib_nodes;   /* library_nodes, for <instance_node> */
};

bool dae_walk(const DaeDoc& doc, const XmlNode& node, const Mat4& parent, Soup& s, uint32_t& next_obj, int depth, std::string& err)
{
    if (depth > 64) { err = "COLLADA: node hierarchy deeper than 64 (cycle?)"; return false; }
    const Mat4 M = mat_mul(parent, dae_node_transform(node));
    for (const XmlNode& c : node.kids) {
        if (c.name == "instance_geometry") {
            const std::string* url = c.attr("url");
            if (!url) continue;
            const std::string id = strip_hash(*url);
            for (auto& g : doc.geoms) {
                if (g.first != id) continue;
                if (g.second.t.empty()) break;
                const uint32_t base = (uint32_t)(s.v.size() / 3);
                for (size_t i = 0; i + 2 < g.second.v.size(); i += 3) {
                    const double x = g.second.v[i], y = g.second.v[i + 1], z = g.second.v[i + 2];
                    s.v.push_back((float)(M.m[0] * x + M.m[1] * y + M.m[2] * z + M.m[3]));
                    s.v.push_back((float)(M.m[4] * x + M.m[5] * y + M.m[6] * z + M.m[7]));
                    s.v.push_back((float)(M.m[8] * x + M.m[9] * y + M.m[10] * z + M.m[11]));
                }
                for (uint32_t i : g.second.t) s.t.push_back(base + i);
                s.o.insert(s.o.end(), g.second.t.size() / 3, next_obj);
                next_obj++;                              /* one object id per mesh instance, in scene-graph order */
                break;
            }
        } else if (c.name == "instance_node") {
            const std::string* url = c.attr("url");
            if (!url) continue;
            const std::string id = strip_hash(*url);
            for (auto& ln : doc.lib_nodes) if (ln.first == id && !dae_walk(doc, *ln.second, M, s, next_obj, depth + 1, err)) return false;
        } else if (c.name == "node") {
            if (!dae_walk(doc, c, M, s, next_obj, depth + 1, err)) return false;
        }
    }
    return true;
}

bool load_dae(const std::string& path, Soup& s, std::string& err)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    const std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    XmlReader rd(data);
    XmlNode root;
    if (!rd.parse(root)) { err = rd.err; return false; }
    if (root.name != "COLLADA") { err = "not a COLLADA document (root element <" + root.name + ">)"; return false; }
    DaeDoc doc;
    for (const XmlNode& lib : root.kids) {
        if (lib.name == "library_geometries") {
            for (const XmlNode& g : lib.kids) {
                if (g.name != "geometry") continue;
                const std::string* id = g.attr("id");
                doc.geoms.emplace_back(id ? *id : std::string(), DaeGeom());
                if (!dae_geometry(g, doc.geoms.back().second, err)) return false;
            }
        } else if (lib.name == "library_nodes") {
            for (const XmlNode& n : lib.kids) if (n.name == "node") if (const std::string* id = n.attr("id")) doc.lib_nodes.emplace_back(*id, &n);
        }
    }
    /* the instantiated scene: <scene><instance_visual_scene url>, else the first visual scene */
    const XmlNode* vs = nullptr;
    std::string want;
    if (const XmlNode* sc = root.child("scene")) if (const XmlNode* iv = sc->child("instance_visual_scene")) if (const std::string* u = iv->attr("url")) want = strip_hash(*u);
    for (const XmlNode& lib : root.kids) {
        if (lib.name != "library_visual_scenes") continue;
        for (const XmlNode& v : lib.kids) {
            if (v.name != "visual_scene") continue;
            const std::string* id = v.attr("id");
            if (!vs || (id && *id == want)) vs = &v;
        }
    }
    /* <asset><up_axis>: the file's axes are KEPT by default (what assimp does under AI_CONFIG_IMPORT_COLLADA_IGNORE_UP_DIRECTION,
     * the setting a ROS map loader needs: map frames are Z-up and Blender exports Z_UP). RR_DAE_UP_AXIS=assimp applies assimp's
     * default instead, the root transform that turns the document into a Y-up scene (ColladaLoader: Z_UP (x,y,z) -> (x,z,-y),
     * X_UP (x,y,z) -> (-y,x,z)). */
    Mat4 root_m = mat_identity();
    if (const XmlNode* asset = root.child("asset")) if (const XmlNode* up = asset->child("up_axis")) {
        std::string a = up->text;
        a.erase(std::remove_if(a.begin(), a.end(), [](unsigned char ch) { return std::isspace(ch); }), a.end());
        if (a != "X_UP" && a != "Y_UP" && a != "Z_UP") { err = "COLLADA: unknown <up_axis> '" + a + "'"; return false; }
        const char* mode = getenv("RR_DAE_UP_AXIS");
        if (mode && std::string(mode) == "assimp") {
            if (a == "Z_UP") { root_m = Mat4{}; root_m.m[0] = 1; root_m.m[6] = 1; root_m.m[9] = -1; root_m.m[15] = 1; }
            else if (a == "X_UP") { root_m = Mat4{}; root_m.m[1] = -1; root_m.m[4] = 1; root_m.m[10] = 1; root_m.m[15] = 1; }
        }
    }
    uint32_t next_obj = 0;
    if (vs) {
        for (const XmlNode& n : vs->kids) if (n.name == "node" && !dae_walk(doc, n, root_m, s, next_obj, 0, err)) return false;
    }
    if (s.t.empty()) {                                   /* no scene graph: every geometry once, untransformed */
        for (auto& g : doc.geoms) {
            if (g.second.t.empty()) continue;
            const uint32_t base = (uint32_t)(s.v.size() / 3);
            s.v.insert(s.v.end(), g.second.v.begin(), g.second.v.end());
            for (uint32_t i : g.second.t) s.t.push_back(base + i);
            s.o.insert(s.o.end(), g.second.t.size() / 3, next_obj++);
        }
    }
    s.n_objects = next_obj;
    return true;
}

void set_err(char* err, size_t cap, const std::string& msg)
{
    if (err && cap) { snprintf(err, cap, "%s", msg.c_str()); }
}

}  // namespace

extern "C" {

int rr_mesh_load(const char* path, rr_mesh* out, char* err, size_t err_cap) try
{
    if (!path || !out) { set_err(err, err_cap, "rr_mesh_load: NULL argument"); return RR_ERR_INVALID_ARGUMENT; }
    memset(out, 0, sizeof(*out));
    Soup s; std::string e;
    const std::string p(path);
    bool ok;
    if (ends_with_ci(p, ".ply")) ok = load_ply(p, s, e);
    else if (ends_with_ci(p, ".obj")) ok = load_obj(p, s, e);
    else if (ends_with_ci(p, ".dae")) ok = load_dae(p, s, e);
    else { set_err(err, err_cap, "rr_mesh_load: unsupported mesh format (supported: .ply, .obj, .dae): " + p); return RR_ERR_INVALID_ARGUMENT; }
    if (!ok) { set_err(err, err_cap, e); return RR_ERR_INVALID_ARGUMENT; }
    if (s.t.empty()) { set_err(err, err_cap, "mesh file holds no faces: " + p); return RR_ERR_INVALID_ARGUMENT; }
    out->n_verts = s.v.size() / 3; out->n_tris = s.t.size() / 3; out->n_objects = s.n_objects;
    out->verts_xyz = (float*)malloc(s.v.size() * sizeof(float));
    out->tri_idx = (uint32_t*)malloc(s.t.size() * sizeof(uint32_t));
    out->tri_object_id = (uint32_t*)malloc(s.o.size() * sizeof(uint32_t));
    if (!out->verts_xyz || !out->tri_idx || !out->tri_object_id) {
        rr_mesh_free(out);
        set_err(err, err_cap, "rr_mesh_load: out of memory");
        return RR_ERR_OUT_OF_MEMORY;
    }
    memcpy(out->verts_xyz, s.v.data(), s.v.size() * sizeof(float));
    memcpy(out->tri_idx, s.t.data(), s.t.size() * sizeof(uint32_t));
    memcpy(out->tri_object_id, s.o.data(), s.o.size() * sizeof(uint32_t));
    return RR_OK;
}
catch (const std::bad_alloc&) { if (out) rr_mesh_free(out); set_err(err, err_cap, "rr_mesh_load: out of memory"); return RR_ERR_OUT_OF_MEMORY; }
catch (const std::length_error&) { if (out) rr_mesh_free(out); set_err(err, err_cap, "rr_mesh_load: a size declared in the file is absurd"); return RR_ERR_OUT_OF_MEMORY; }
catch (const std::exception& ex) { if (out) rr_mesh_free(out); set_err(err, err_cap, std::string("rr_mesh_load: ") + ex.what()); return RR_ERR_INVALID_ARGUMENT; }
catch (...) { if (out) rr_mesh_free(out); set_err(err, err_cap, "rr_mesh_load: unknown exception"); return RR_ERR_INVALID_ARGUMENT; }

void rr_mesh_free(rr_mesh* m)
{
    if (!m) return;
    free(m->verts_xyz); free(m->tri_idx); free(m->tri_object_id);
    memset(m, 0, sizeof(*m));
}

}  // extern "C"
