// b2r_obj.cpp -- native Wavefront OBJ tokenizer, host only.
//
// Produces the very arrays `Model.load_model` builds line by line in Python (reference: obj/core.py:257-318):
//   vertices float32 (V,4)  "v x y z [w]"       w = 1 appended when absent            (core.py:281-285)
//   uv       float32 (T,3)  "vt u v [w]"        0 appended when absent                (core.py:304-309)
//   normals  float32 (N,3)  "vn x y z"                                                (core.py:301-303)
//   faces    int32 (F,3,4)  "f a/b/c ..."       fan triangulation (core.py:72-74), missing index -> -1, 4th column =
//                                               material slot + 1, then `where(x > 0, x - 1, x)` (core.py:313)
// plus the `usemtl` names in first-use order and the `mtllib` file names (parsed by the Python side).
// Numbers go through strtod and are then narrowed to float, which is what NumPy does with a list of strings.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b2r.h"

namespace {
struct Cursor {
    const char* p;
    const char* end;
    void skip_blanks() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; }
    bool at_eol() const { return p >= end || *p == '\n'; }
    void next_line() { while (p < end && *p != '\n') ++p; if (p < end) ++p; }
    std::string word() { skip_blanks(); const char* s = p; while (p < end && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') ++p; return std::string(s, p); }
};
}  // namespace

// Return codes: 0 ok, 1 bad arguments, 2 file cannot be read (FileNotFoundError / OSError on the Python side),
// 3 malformed number in a face statement (the reference raises ValueError from np.array(..., dtype=int32),
// core.py:72-74), 4 out of memory.  No C++ exception crosses the C boundary.
static int obj_load_impl(const char* path, b2r_obj* out);
extern "C" int b2r_obj_load(const char* path, b2r_obj* out) {
    if (!path || !out) return 1;
    std::memset(out, 0, sizeof(*out));
    try {
        return obj_load_impl(path, out);
    } catch (...) {
        b2r_obj_free(out);
        return 4;
    }
}

static int obj_load_impl(const char* path, b2r_obj* out) {
    FILE* fh = std::fopen(path, "rb");
    if (!fh) return 2;
    if (std::fseek(fh, 0, SEEK_END) != 0) { std::fclose(fh); return 2; }
    const long size = std::ftell(fh);
    if (size < 0 || std::fseek(fh, 0, SEEK_SET) != 0) { std::fclose(fh); return 2; }  // e.g. a directory
    std::string text((size_t)size, '\0');
    if (size > 0 && std::fread(&text[0], 1, (size_t)size, fh) != (size_t)size) { std::fclose(fh); return 2; }
    std::fclose(fh);

    std::vector<float> v, vt, vn;
    std::vector<int32_t> faces;
    std::vector<std::string> slots{"default"};
    std::string mtllibs;
    int current = 0;
    Cursor c{text.data(), text.data() + text.size()};
    std::vector<int32_t> corner;  // 4 ints per polygon corner
    while (c.p < c.end) {
        const std::string head = c.word();
        if (head == "v" || head == "vt" || head == "vn") {
            double val[4];
            int n = 0;
            for (;;) {
                c.skip_blanks();
                if (c.at_eol() || n == 4) break;
                char* stop = nullptr;
                val[n] = std::strtod(c.p, &stop);
                if (stop == c.p) break;
                c.p = stop;
                ++n;
            }
            if (head == "v") {
                if (n == 3) val[n++] = 1.0;
                for (int i = 0; i < 4; ++i) v.push_back(i < n ? (float)val[i] : 0.0f);
            } else if (head == "vt") {
                if (n == 2) val[n++] = 0.0;
                for (int i = 0; i < 3; ++i) vt.push_back(i < n ? (float)val[i] : 0.0f);
            } else {
                for (int i = 0; i < 3; ++i) vn.push_back(i < n ? (float)val[i] : 0.0f);
            }
        } else if (head == "f") {
            corner.clear();
            for (;;) {
                c.skip_blanks();
                if (c.at_eol()) break;
                int32_t idx[3] = {-1, -1, -1};
                for (int k = 0; k < 3; ++k) {
                    if (c.p < c.end && *c.p != '/' && *c.p != ' ' && *c.p != '\t' && *c.p != '\r' && *c.p != '\n') {
                        char* stop = nullptr;
                        idx[k] = (int32_t)std::strtol(c.p, &stop, 10);
                        if (stop == c.p) return 3;  // not a number: int('x') raises ValueError in the reference
                        c.p = stop;
                    }
                    if (c.p < c.end && *c.p == '/') ++c.p; else break;
                }
                // a corner token must end here ("1x", "1/2/3/4" are malformed in the reference as well)
                if (c.p < c.end && *c.p != ' ' && *c.p != '\t' && *c.p != '\r' && *c.p != '\n') return 3;
                for (int k = 0; k < 3; ++k) corner.push_back(idx[k]);
                corner.push_back(current + 1);
            }
            const int nc = (int)corner.size() / 4;
            for (int k = 1; k + 1 < nc; ++k) {  // fan: (0, k, k+1)
                const int tri[3] = {0, k, k + 1};
                for (int t = 0; t < 3; ++t)
                    for (int j = 0; j < 4; ++j) {
                        const int32_t x = corner[(size_t)tri[t] * 4 + j];
                        faces.push_back(x > 0 ? x - 1 : x);
                    }
            }
        } else if (head == "usemtl") {
            const std::string name = c.word();
            current = -1;
            for (size_t i = 0; i < slots.size(); ++i) if (slots[i] == name) current = (int)i;
            if (current < 0) { slots.push_back(name); current = (int)slots.size() - 1; }
        } else if (head == "mtllib") {
            const std::string name = c.word();
            mtllibs += name;
            mtllibs += '\n';
        }
        c.next_line();
    }
    auto dup = [](const void* src, size_t bytes) { void* p = std::malloc(bytes ? bytes : 1); if (p && bytes) std::memcpy(p, src, bytes); return p; };
    std::string names;
    for (const std::string& s : slots) { names += s; names += '\n'; }
    out->vertices = (float*)dup(v.data(), v.size() * sizeof(float)); out->n_vertices = (int32_t)(v.size() / 4);
    out->uv = (float*)dup(vt.data(), vt.size() * sizeof(float)); out->n_uv = (int32_t)(vt.size() / 3);
    out->normals = (float*)dup(vn.data(), vn.size() * sizeof(float)); out->n_normals = (int32_t)(vn.size() / 3);
    out->faces = (int32_t*)dup(faces.data(), faces.size() * sizeof(int32_t)); out->n_faces = (int32_t)(faces.size() / 12);
    out->slot_names = (char*)dup(names.c_str(), names.size() + 1);
    out->mtllibs = (char*)dup(mtllibs.c_str(), mtllibs.size() + 1);
    return 0;
}

extern "C" void b2r_obj_free(b2r_obj* o) {
    if (!o) return;
    std::free(o->vertices); std::free(o->uv); std::free(o->normals); std::free(o->faces);
    std::free(o->slot_names); std::free(o->mtllibs);
    std::memset(o, 0, sizeof(*o));
}
