// JQ4 / JQ8 checkpoint I/O: the safetensors container with Jlama's non-standard dtype strings ("Q4", "I8" + "<name>.qb"
// f32 block-scale tensors), the model.safetensors.index.json shard map, the offline quantiser and the loader that binds a
// checkpoint to a device-resident model.  Reference:
//   core/safetensors/SafeTensorSupport.java:54-102 (readTensorInfoMap), :215-332 (quantizeModel)
//   core/safetensors/Weights.java:49-66 (majority dtype ignores .qb), :99-179 (load offsets, Q4/I8 + .qb pairing)
//   core/safetensors/SafeTensorIndex.java:59-236 (index.json, mmap splits), core/safetensors/TensorInfo.java:26-46
//   core/tensor/AbstractTensor.java:278-298 (which tensors quantise), :300-312 (save)
// The reference maps files in <= 2 GiB pieces (Integer.MAX_VALUE mmap limit, SafeTensorIndex.java:121-236) and re-assembles
// tensors that straddle a piece; here a file is ONE 64-bit mapping, so tensors of any size are read in place.
#include "jl_common.cuh"

#include <algorithm>
#include <fcntl.h>
#include <map>
#include <memory>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

// ---- a small JSON reader (objects, arrays, strings, numbers, literals): enough for safetensors headers, index.json, config.json ----
struct JVal {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    bool b = false;
    double num = 0;
    long long inum = 0;
    bool is_int = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj; // insertion order kept
    const JVal *get(const char *key) const {
        for (auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};
struct JParser {
    const char *p, *end;
    std::string err;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool fail(const char *m) {
        if (err.empty()) err = m;
        return false;
    }
    bool parse_string(std::string &out) {
        if (p >= end || *p != '"') return fail("expected string");
        p++;
        out.clear();
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (end - p < 5) return fail("bad \\u escape");
                        unsigned cp = 0;
                        for (int i = 1; i <= 4; i++) {
                            const char c = p[i];
                            cp = cp * 16 + (c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : (c >= 'A' && c <= 'F' ? c - 'A' + 10 : 0)));
                        }
                        p += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) out += (char)(0xC0 | (cp >> 6)), out += (char)(0x80 | (cp & 0x3F));
                        else out += (char)(0xE0 | (cp >> 12)), out += (char)(0x80 | ((cp >> 6) & 0x3F)), out += (char)(0x80 | (cp & 0x3F));
                    } break;
                    default: out += *p;
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (p >= end) return fail("unterminated string");
        p++;
        return true;
    }
    bool parse(JVal &v, int depth = 0) {
        if (depth > 64) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '{') {
            v.kind = JVal::OBJ;
            p++;
            ws();
            if (p < end && *p == '}') return p++, true;
            for (;;) {
                ws();
                std::string k;
                if (!parse_string(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                p++;
                JVal c;
                if (!parse(c, depth + 1)) return false;
                v.obj.emplace_back(std::move(k), std::move(c));
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == '}') return p++, true;
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JVal::ARR;
            p++;
            ws();
            if (p < end && *p == ']') return p++, true;
            for (;;) {
                JVal c;
                if (!parse(c, depth + 1)) return false;
                v.arr.push_back(std::move(c));
                ws();
                if (p < end && *p == ',') {
                    p++;
                    continue;
                }
                if (p < end && *p == ']') return p++, true;
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') {
            v.kind = JVal::STR;
            return parse_string(v.str);
        }
        if (!strncmp(p, "true", std::min<size_t>(4, end - p)) && end - p >= 4) return v.kind = JVal::BOOL, v.b = true, p += 4, true;
        if (!strncmp(p, "false", std::min<size_t>(5, end - p)) && end - p >= 5) return v.kind = JVal::BOOL, v.b = false, p += 5, true;
        if (!strncmp(p, "null", std::min<size_t>(4, end - p)) && end - p >= 4) return v.kind = JVal::NUL, p += 4, true;
        // number
        const char *s = p;
        bool isint = true;
        if (p < end && (*p == '-' || *p == '+')) p++;
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
            if (*p == '.' || *p == 'e' || *p == 'E') isint = false;
            p++;
        }
        if (p == s) return fail("unexpected character");
        std::string t(s, p);
        v.kind = JVal::NUM;
        v.num = strtod(t.c_str(), nullptr);
        v.is_int = isint;
        v.inum = isint ? strtoll(t.c_str(), nullptr, 10) : (long long)v.num;
        return true;
    }
};

static bool json_parse(const char *s, size_t n, JVal &out, std::string &err) {
    JParser ps{s, s + n, ""};
    if (!ps.parse(out)) {
        err = ps.err;
        return false;
    }
    return true;
}

// ---- container -----------------------------------------------------------------------------------------------------------
#define JL_ST_F16 4 // read-only: converted to F32 when bound to a model (Weights.java:137-152 does the same for F32-majority models)

struct StFile {
    std::string path;
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t size = 0, data0 = 0; // data0 = 8 + header length
};
struct StTensor {
    std::string name;
    int dtype = -1;
    int ndim = 0;
    int64_t shape[4] = {0, 0, 0, 0};
    int file = 0;
    uint64_t off0 = 0, off1 = 0; // relative to the file's data section
};
struct jl_st {
    std::vector<StFile> files;
    std::vector<StTensor> tensors; // sorted by (file, offset) like readTensorInfoMap sorts by data_offsets
    std::map<std::string, int> by_name;
    std::map<std::string, std::string> metadata;
    std::string dir;
    std::string error;
};

static thread_local std::string g_st_error;
extern "C" const char *jl_st_last_error(void) { return g_st_error.c_str(); }
static int st_fail(const std::string &m, int code = JL_ERR_INVALID) {
    g_st_error = m;
    return code;
}

static int dtype_from_string(const std::string &s) {
    if (s == "F32") return JL_F32;
    if (s == "BF16") return JL_BF16;
    if (s == "Q4") return JL_Q4;
    if (s == "I8") return JL_I8;
    if (s == "F16") return JL_ST_F16;
    return -1;
}
static const char *dtype_to_string(int d) {
    switch (d) {
        case JL_F32: return "F32";
        case JL_BF16: return "BF16";
        case JL_Q4: return "Q4";
        case JL_I8: return "I8";
        case JL_ST_F16: return "F16";
    }
    return "?";
}

static int st_map_file(jl_st *st, const std::string &path) {
    StFile f;
    f.path = path;
    f.fd = open(path.c_str(), O_RDONLY);
    if (f.fd < 0) return st_fail("cannot open " + path);
    struct stat sb;
    if (fstat(f.fd, &sb) != 0 || sb.st_size < 8) {
        close(f.fd);
        return st_fail(path + ": not a safetensors file");
    }
    f.size = (size_t)sb.st_size;
    void *m = mmap(nullptr, f.size, PROT_READ, MAP_PRIVATE, f.fd, 0);
    if (m == MAP_FAILED) {
        close(f.fd);
        return st_fail("mmap failed for " + path);
    }
    f.map = (const uint8_t *)m;
    // SafeTensorSupport.readTensorInfoMap (:54-77): 8-byte little-endian header length, < 0 or > 1 GiB rejected
    int64_t hlen;
    memcpy(&hlen, f.map, 8);
    if (hlen < 0) {
        munmap(m, f.size), close(f.fd);
        return st_fail("Header length cannot be negative: " + std::to_string(hlen));
    }
    if (hlen > (1LL << 30) || (uint64_t)hlen + 8 > f.size) {
        munmap(m, f.size), close(f.fd);
        return st_fail("Header length " + std::to_string(hlen) + " exceeds the maximum allowed length");
    }
    f.data0 = 8 + (size_t)hlen;
    JVal root;
    std::string err;
    if (!json_parse((const char *)f.map + 8, (size_t)hlen, root, err) || root.kind != JVal::OBJ) {
        munmap(m, f.size), close(f.fd);
        return st_fail(path + ": bad header json: " + err);
    }
    const int fi = (int)st->files.size();
    for (auto &kv : root.obj) {
        if (strcasecmp(kv.first.c_str(), "__metadata__") == 0) { // equalsIgnoreCase (:84)
            for (auto &m2 : kv.second.obj)
                if (m2.second.kind == JVal::STR) st->metadata[m2.first] = m2.second.str;
            continue;
        }
        const JVal *dt = kv.second.get("dtype"), *sh = kv.second.get("shape"), *off = kv.second.get("data_offsets");
        auto bad = [&](const std::string &why) {
            munmap(m, f.size), close(f.fd);
            return st_fail(path + ": " + why + " (tensor " + kv.first + ")");
        };
        if (kv.second.kind != JVal::OBJ || !dt || dt->kind != JVal::STR || !sh || sh->kind != JVal::ARR || !off || off->kind != JVal::ARR ||
            off->arr.size() != 2 || sh->arr.size() > 4)
            return bad("malformed tensor entry");
        StTensor t;
        t.name = kv.first;
        t.dtype = dtype_from_string(dt->str);
        t.ndim = (int)sh->arr.size();
        // everything below is read through these numbers: they must be non-negative integers, the byte range must lie inside the
        // file's data section and, for the dtypes this library reads, hold exactly the elements the shape promises
        uint64_t elems = 1;
        for (int i = 0; i < t.ndim; i++) {
            const JVal &dv = sh->arr[(size_t)i];
            if (dv.kind != JVal::NUM || !dv.is_int || dv.inum < 0) return bad("shape entries must be non-negative integers");
            t.shape[i] = dv.inum;
            if (dv.inum != 0 && elems > (1ULL << 62) / (uint64_t)dv.inum) return bad("shape overflows");
            elems *= (uint64_t)dv.inum;
        }
        for (int i = 0; i < 2; i++)
            if (off->arr[(size_t)i].kind != JVal::NUM || !off->arr[(size_t)i].is_int || off->arr[(size_t)i].inum < 0)
                return bad("data_offsets must be non-negative integers");
        t.file = fi;
        t.off0 = (uint64_t)off->arr[0].inum, t.off1 = (uint64_t)off->arr[1].inum;
        if (t.off1 < t.off0 || t.off1 > f.size - f.data0) return bad("data_offsets exceed the file");
        if (t.dtype >= 0) {
            if (t.dtype == JL_Q4 && (elems & 1)) return bad("Q4 tensor with an odd element count");
            const uint64_t want = t.dtype == JL_F32 ? elems * 4 : (t.dtype == JL_Q4 ? elems / 2 : (t.dtype == JL_I8 ? elems : elems * 2));
            if (want != t.off1 - t.off0) return bad("byte range does not match dtype and shape");
        }
        st->tensors.push_back(t);
    }
    st->files.push_back(f);
    return JL_OK;
}

static bool file_exists(const std::string &p) {
    struct stat sb;
    return stat(p.c_str(), &sb) == 0;
}
static bool is_dir(const std::string &p) {
    struct stat sb;
    return stat(p.c_str(), &sb) == 0 && S_ISDIR(sb.st_mode);
}
static bool read_text(const std::string &p, std::string &out) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? n : 0);
    const bool ok = n <= 0 || fread(&out[0], 1, n, f) == (size_t)n;
    fclose(f);
    return ok;
}

extern "C" int jl_st_close(jl_st *st) {
    if (!st) return JL_ERR_INVALID;
    for (auto &f : st->files) {
        if (f.map) munmap((void *)f.map, f.size);
        if (f.fd >= 0) close(f.fd);
    }
    delete st;
    return JL_OK;
}

// `path`: a .safetensors file, or a model directory holding model.safetensors or model.safetensors.index.json (+ shards)
// (SafeTensorSupport.loadWeights :104-121, SafeTensorIndex.loadWithWeights :59-70)
extern "C" int jl_st_open(const char *path, jl_st **out) {
    if (!path || !out) return JL_ERR_INVALID;
    *out = nullptr;
    std::unique_ptr<jl_st> st(new jl_st());
    std::string p(path);
    int rc = JL_OK;
    if (is_dir(p)) {
        st->dir = p;
        const std::string idx = p + "/model.safetensors.index.json", single = p + "/model.safetensors";
        if (file_exists(idx)) {
            std::string txt, err;
            JVal root;
            if (!read_text(idx, txt) || !json_parse(txt.data(), txt.size(), root, err)) return st_fail("bad index json " + idx + ": " + err);
            const JVal *wm = root.get("weight_map");
            if (!wm || wm->kind != JVal::OBJ) return st_fail(idx + ": no weight_map");
            std::vector<std::string> shards;
            for (auto &kv : wm->obj)
                if (kv.second.kind == JVal::STR && std::find(shards.begin(), shards.end(), kv.second.str) == shards.end()) shards.push_back(kv.second.str);
            for (auto &s : shards) {
                // shard files live next to the index (SafeTensorIndex.java:59-70): no absolute paths, no way out of the directory
                if (s.empty() || s[0] == '/' || s.find("..") != std::string::npos) {
                    rc = st_fail(idx + ": bad shard file name " + s);
                    break;
                }
                if ((rc = st_map_file(st.get(), p + "/" + s)) != JL_OK) break;
            }
        } else if (file_exists(single)) {
            rc = st_map_file(st.get(), single);
        } else {
            return st_fail("no model.safetensors(.index.json) under " + p);
        }
    } else {
        rc = st_map_file(st.get(), p);
    }
    if (rc != JL_OK) {
        jl_st_close(st.release());
        return rc;
    }
    std::stable_sort(st->tensors.begin(), st->tensors.end(), [](const StTensor &a, const StTensor &b) {
        return a.file != b.file ? a.file < b.file : a.off0 < b.off0; // TensorInfo.compareTo (:86-90)
    });
    for (size_t i = 0; i < st->tensors.size(); i++) st->by_name[st->tensors[i].name] = (int)i;
    *out = st.release();
    return JL_OK;
}

extern "C" int jl_st_count(jl_st *st) { return st ? (int)st->tensors.size() : JL_ERR_INVALID; }
extern "C" int jl_st_find(jl_st *st, const char *name) {
    if (!st || !name) return JL_ERR_INVALID;
    auto it = st->by_name.find(name);
    return it == st->by_name.end() ? -1 : it->second;
}
extern "C" int jl_st_info(jl_st *st, int i, const char **name, int *dtype, int *ndim, int64_t *shape4, int64_t *nbytes) {
    if (!st || i < 0 || i >= (int)st->tensors.size()) return JL_ERR_INVALID;
    const StTensor &t = st->tensors[i];
    if (name) *name = t.name.c_str();
    if (dtype) *dtype = t.dtype;
    if (ndim) *ndim = t.ndim;
    if (shape4)
        for (int k = 0; k < 4; k++) shape4[k] = t.shape[k];
    if (nbytes) *nbytes = (int64_t)(t.off1 - t.off0);
    return JL_OK;
}
extern "C" const void *jl_st_data(jl_st *st, int i) {
    if (!st || i < 0 || i >= (int)st->tensors.size()) return nullptr;
    const StTensor &t = st->tensors[i];
    return st->files[t.file].map + st->files[t.file].data0 + t.off0;
}
extern "C" const char *jl_st_metadata(jl_st *st, const char *key) {
    if (!st || !key) return nullptr;
    auto it = st->metadata.find(key);
    return it == st->metadata.end() ? nullptr : it->second.c_str();
}
// Weights.findDType (:49-66): the most frequent dtype, ".qb" scale tensors not counted; F16 counts as F32
extern "C" int jl_st_majority_dtype(jl_st *st) {
    if (!st) return JL_ERR_INVALID;
    int counts[8] = {0};
    for (auto &t : st->tensors) {
        const size_t n = t.name.size();
        if (n >= 3 && t.name.compare(n - 3, 3, ".qb") == 0) continue;
        if (t.dtype >= 0 && t.dtype < 8) counts[t.dtype]++;
    }
    int best = -1, mx = 0;
    for (int d = 0; d < 8; d++)
        if (counts[d] > mx) mx = counts[d], best = d;
    return best == JL_ST_F16 ? JL_F32 : best;
}

static std::string json_escape(const std::string &s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\', o += c;
        else if (c == '\n') o += "\\n";
        else if ((unsigned char)c < 0x20) {
            char b[8];
            snprintf(b, sizeof b, "\\u%04x", c);
            o += b;
        } else o += c;
    }
    return o;
}

// Write one safetensors file: 8-byte LE header length, JSON header, raw tensor bytes in the order given
// (SafeTensorSupport.quantizeModel :306-329 / AbstractTensor.save :300-312: offsets are relative to the data section).
extern "C" int jl_st_write(const char *path, int n, const char *const *names, const int *dtypes, const int *ndims, const int64_t *shapes4,
                           const void *const *data, const int64_t *nbytes, int n_meta, const char *const *meta_kv) {
    if (!path || n < 0 || (n && (!names || !dtypes || !ndims || !shapes4 || !data || !nbytes))) return JL_ERR_INVALID;
    for (int i = 0; i < n; i++) {
        if (!names[i] || ndims[i] < 0 || ndims[i] > 4 || nbytes[i] < 0 || (nbytes[i] > 0 && !data[i]))
            return st_fail("st_write: tensor " + std::to_string(i) + " has a null name / data pointer, more than 4 dims or a negative size");
        for (int k = 0; k < ndims[i]; k++)
            if (shapes4[(size_t)i * 4 + k] < 0) return st_fail(std::string("st_write: negative dimension in ") + names[i]);
    }
    if (n_meta < 0 || (n_meta > 0 && !meta_kv)) return st_fail("st_write: metadata count without metadata");
    for (int i = 0; i < 2 * n_meta; i++)
        if (!meta_kv[i]) return st_fail("st_write: null metadata string");
    std::string h = "{";
    uint64_t off = 0;
    for (int i = 0; i < n; i++) {
        if (i) h += ",";
        h += "\"" + json_escape(names[i]) + "\":{\"dtype\":\"" + dtype_to_string(dtypes[i]) + "\",\"shape\":[";
        for (int k = 0; k < ndims[i]; k++) h += (k ? "," : "") + std::to_string((long long)shapes4[(size_t)i * 4 + k]);
        h += "],\"data_offsets\":[" + std::to_string(off) + "," + std::to_string(off + (uint64_t)nbytes[i]) + "]}";
        off += (uint64_t)nbytes[i];
    }
    if (n_meta > 0 && meta_kv) {
        h += std::string(n ? "," : "") + "\"__metadata__\":{";
        for (int i = 0; i < n_meta; i++) h += std::string(i ? "," : "") + "\"" + json_escape(meta_kv[2 * i]) + "\":\"" + json_escape(meta_kv[2 * i + 1]) + "\"";
        h += "}";
    }
    h += "}";
    FILE *f = fopen(path, "wb");
    if (!f) return st_fail(std::string("cannot create ") + path);
    const int64_t hlen = (int64_t)h.size();
    bool ok = fwrite(&hlen, 8, 1, f) == 1 && fwrite(h.data(), 1, h.size(), f) == h.size();
    for (int i = 0; ok && i < n; i++) ok = nbytes[i] == 0 || fwrite(data[i], 1, (size_t)nbytes[i], f) == (size_t)nbytes[i];
    ok = fclose(f) == 0 && ok;
    return ok ? JL_OK : st_fail(std::string("write failed: ") + path);
}

// ---- offline quantiser: SafeTensorSupport.quantizeModel (:215-332) with the block quantisers running on the GPU -----------------
// `src` points into the file mapping: the reference writes the data section right after an unpadded header (SafeTensorSupport.java:306-311),
// so 2- and 4-byte elements may sit at odd addresses; they are read bytewise.
static inline uint16_t load_u16(const void *base, size_t i) {
    uint16_t v;
    memcpy(&v, (const uint8_t *)base + 2 * i, 2);
    return v;
}
static void f16_to_f32(const void *src, float *dst, size_t n) {
    for (size_t i = 0; i < n; i++) {
        const uint32_t h = load_u16(src, i), s = (h & 0x8000u) << 16, e = (h >> 10) & 0x1F, m = h & 0x3FF;
        uint32_t u;
        if (e == 0) {
            if (m == 0) u = s;
            else { // subnormal
                int sh = 0;
                uint32_t mm = m;
                while (!(mm & 0x400)) mm <<= 1, sh++;
                u = s | ((uint32_t)(127 - 15 - sh + 1) << 23) | ((mm & 0x3FF) << 13);
            }
        } else if (e == 31) u = s | 0x7F800000u | (m << 13);
        else u = s | ((e + 112) << 23) | (m << 13);
        memcpy(&dst[i], &u, 4);
    }
}
static void tensor_to_f32(const StTensor &t, const void *src, std::vector<float> &out) {
    size_t n = 1;
    for (int k = 0; k < t.ndim; k++) n *= (size_t)t.shape[k];
    out.resize(n);
    if (t.dtype == JL_F32) memcpy(out.data(), src, n * 4);
    else if (t.dtype == JL_BF16) {
        for (size_t i = 0; i < n; i++) {
            const uint32_t u = (uint32_t)load_u16(src, i) << 16; // FloatConversions.bFloat16ToFloat32 (:31-33)
            memcpy(&out[i], &u, 4);
        }
    } else f16_to_f32(src, out.data(), n);
}
static std::vector<std::string> split_csv(const char *s) {
    std::vector<std::string> v;
    if (!s) return v;
    std::string cur;
    for (const char *p = s;; p++) {
        if (*p == ',' || *p == 0) {
            if (!cur.empty()) v.push_back(cur);
            cur.clear();
            if (!*p) break;
        } else cur += *p;
    }
    return v;
}
static bool copy_file(const std::string &a, const std::string &b) {
    std::string t;
    if (!read_text(a, t)) return false;
    FILE *f = fopen(b.c_str(), "wb");
    if (!f) return false;
    const bool ok = t.empty() || fwrite(t.data(), 1, t.size(), f) == t.size();
    return fclose(f) == 0 && ok;
}

extern "C" int jl_quantize_q8_weights(jl_ctx *ctx, const float *x, int64_t rows, int64_t cols, int8_t *q, float *scales);

// qtype: JL_Q4 or JL_I8.  skip_csv: substrings of tensor names that keep their dtype (default "norm", QuantizeCommand.java:36-38);
// drop_csv: name prefixes that are not written.  Writes dst_dir/model.safetensors (+ config.json, tokenizer files when present).
extern "C" int jl_quantize_model(jl_ctx *ctx, const char *src_dir, const char *dst_dir, int qtype, const char *skip_csv, const char *drop_csv) {
    if (!ctx || !src_dir || !dst_dir || (qtype != JL_Q4 && qtype != JL_I8)) return JL_ERR_INVALID;
    jl_st *st = nullptr;
    int rc = jl_st_open(src_dir, &st);
    if (rc != JL_OK) return jl_set_error(ctx, rc, "quantize_model: %s", g_st_error.c_str());
    const std::vector<std::string> skip = skip_csv ? split_csv(skip_csv) : std::vector<std::string>{"norm"}, drop = split_csv(drop_csv);
    struct Out {
        std::string name;
        int dtype, ndim;
        int64_t shape[4];
        std::vector<uint8_t> own; // quantised bytes (empty: points into the mapping)
        const void *ptr;
        int64_t nbytes;
    };
    std::vector<Out> outs;
    outs.reserve(st->tensors.size() * 2);
    for (size_t i = 0; i < st->tensors.size(); i++) {
        const StTensor &t = st->tensors[i];
        bool dropped = false;
        for (auto &d : drop) dropped = dropped || t.name.compare(0, d.size(), d) == 0; // startsWith (:234)
        if (dropped) continue;
        bool skipq = false;
        for (auto &s : skip) skipq = skipq || t.name.find(s) != std::string::npos; // contains (:247)
        const void *src = jl_st_data(st, (int)i);
        // AbstractTensor.quantize (:282-284): first dim 1, same dtype, or a narrower source dtype stay as they are
        const int src_size = t.dtype == JL_F32 ? 4 : ((t.dtype == JL_BF16 || t.dtype == JL_ST_F16) ? 2 : 1);
        const bool keep = skipq || t.ndim != 2 || t.shape[0] == 1 || t.dtype == qtype || src_size < 1 || t.dtype == JL_Q4 || t.dtype == JL_I8 ||
                          (t.shape[1] % 32) != 0;
        Out o;
        o.name = t.name, o.ndim = t.ndim;
        memcpy(o.shape, t.shape, sizeof o.shape);
        if (keep) {
            o.dtype = t.dtype, o.ptr = src, o.nbytes = (int64_t)(t.off1 - t.off0);
            outs.push_back(std::move(o));
            continue;
        }
        std::vector<float> f;
        tensor_to_f32(t, src, f);
        const int64_t rows = t.shape[0], cols = t.shape[1];
        Out sc;
        sc.name = t.name + ".qb", sc.dtype = JL_F32, sc.ndim = 2, sc.shape[0] = rows, sc.shape[1] = cols / 32, sc.shape[2] = sc.shape[3] = 0;
        sc.own.resize((size_t)rows * (cols / 32) * 4);
        if (qtype == JL_Q4) {
            o.own.resize((size_t)rows * cols / 2);
            rc = jl_quantize_q4_weights(ctx, f.data(), rows, cols, o.own.data(), (float *)sc.own.data());
        } else {
            o.own.resize((size_t)rows * cols);
            rc = jl_quantize_q8_weights(ctx, f.data(), rows, cols, (int8_t *)o.own.data(), (float *)sc.own.data());
        }
        if (rc != JL_OK) {
            jl_st_close(st);
            return rc;
        }
        o.dtype = qtype, o.ptr = o.own.data(), o.nbytes = (int64_t)o.own.size();
        sc.ptr = sc.own.data(), sc.nbytes = (int64_t)sc.own.size();
        outs.push_back(std::move(o));
        outs.push_back(std::move(sc));
        outs[outs.size() - 2].ptr = outs[outs.size() - 2].own.data(); // the vectors moved
        outs[outs.size() - 1].ptr = outs[outs.size() - 1].own.data();
    }
    mkdir(dst_dir, 0755);
    std::vector<const char *> names;
    std::vector<int> dts, nds;
    std::vector<int64_t> shp, nb;
    std::vector<const void *> ptrs;
    for (auto &o : outs) {
        names.push_back(o.name.c_str()), dts.push_back(o.dtype), nds.push_back(o.ndim), ptrs.push_back(o.own.empty() ? o.ptr : o.own.data()), nb.push_back(o.nbytes);
        for (int k = 0; k < 4; k++) shp.push_back(o.shape[k]);
    }
    rc = jl_st_write((std::string(dst_dir) + "/model.safetensors").c_str(), (int)outs.size(), names.data(), dts.data(), nds.data(), shp.data(), ptrs.data(),
                     nb.data(), 0, nullptr);
    if (rc == JL_OK && !st->dir.empty())
        for (const char *fn : {"config.json", "tokenizer.json", "tokenizer_config.json", "README.md"})
            if (file_exists(st->dir + "/" + fn)) copy_file(st->dir + "/" + fn, std::string(dst_dir) + "/" + fn);
    jl_st_close(st);
    return rc == JL_OK ? JL_OK : jl_set_error(ctx, rc, "quantize_model: %s", g_st_error.c_str());
}

// ---- config.json -> jl_model_config (safetensors/Config.java:253-287, llama/LlamaConfig.java:27-57) ----------------------------
extern "C" int jl_config_from_json(const char *config_json_path, jl_model_config *cfg) {
    if (!config_json_path || !cfg) return JL_ERR_INVALID;
    std::string txt, err;
    JVal root;
    if (!read_text(config_json_path, txt) || !json_parse(txt.data(), txt.size(), root, err) || root.kind != JVal::OBJ)
        return st_fail(std::string("bad config json ") + config_json_path + ": " + err);
    auto geti = [&](const char *k, long long d) {
        const JVal *v = root.get(k);
        return v && v->kind == JVal::NUM ? v->inum : d;
    };
    auto getd = [&](const char *k, double d) {
        const JVal *v = root.get(k);
        return v && v->kind == JVal::NUM ? v->num : d;
    };
    memset(cfg, 0, sizeof *cfg);
    cfg->context_length = (int)geti("max_position_embeddings", 0);
    cfg->embedding_length = (int)geti("hidden_size", 0);
    cfg->hidden_length = (int)geti("intermediate_size", 0);
    cfg->num_heads = (int)geti("num_attention_heads", 0);
    cfg->num_kv_heads = (int)geti("num_key_value_heads", cfg->num_heads);
    cfg->num_layers = (int)geti("num_hidden_layers", 0);
    cfg->vocab_size = (int)geti("vocab_size", 0);
    cfg->head_size = (int)geti("head_dim", cfg->num_heads ? cfg->embedding_length / cfg->num_heads : 0); // Config.java:254
    cfg->layer_norm_eps = (float)getd("rms_norm_eps", 1e-5);
    cfg->rope_theta = getd("rope_theta", 10000.0);
    cfg->rope_scaling = 1.0;
    if (const JVal *rs = root.get("rope_scaling")) // only rope_type "linear" is honoured (LlamaConfig.java:55-56)
        if (rs->kind == JVal::OBJ) {
            const JVal *ty = rs->get("rope_type"), *fa = rs->get("factor");
            if (ty && ty->kind == JVal::STR && ty->str == "linear" && fa && fa->kind == JVal::NUM) cfg->rope_scaling = fa->num;
        }
    // MixtralConfig (core/model/mixtral/MixtralConfig.java): num_local_experts / num_experts_per_tok
    cfg->num_experts = (int)geti("num_local_experts", 0);
    cfg->experts_per_token = cfg->num_experts > 0 ? (int)geti("num_experts_per_tok", 2) : 0;
    cfg->working_qtype = JL_I8;
    cfg->kv_dtype = JL_F32;
    cfg->tp_size = 1;
    if (cfg->embedding_length <= 0 || cfg->num_heads <= 0 || cfg->num_layers <= 0 || cfg->vocab_size <= 0 || cfg->hidden_length <= 0)
        return st_fail(std::string(config_json_path) + ": missing model dimensions");
    return JL_OK;
}

// ---- bind a checkpoint to a model: LlamaModel.loadInputWeights / loadTransformerBlockWeights / loadOutputWeights ----------------
// (llama/LlamaModel.java:68-156) with the jlama-net shard slices (Weights.getLoadOffsets :99-117 for rows, AbstractTensor.sparsify for
// columns) cut on the host.  Quantised tensors pair with "<name>.qb".  F16 tensors become F32.
struct HostSlice {
    std::vector<uint8_t> data;
    std::vector<float> scales;
};
static int register_slice(jl_ctx *ctx, jl_st *st, const char *name, int64_t row0, int64_t rows, int64_t col0, int64_t cols, int64_t *id_out) {
    const int i = jl_st_find(st, name);
    if (i < 0) return jl_set_error(ctx, JL_ERR_INVALID, "checkpoint has no tensor %s", name);
    const StTensor &t = st->tensors[i];
    if (t.dtype < 0) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "%s: dtype not readable by this library", name);
    if (t.ndim != 1 && t.ndim != 2) return jl_set_error(ctx, JL_ERR_INVALID, "%s: %d-dimensional tensor where a matrix or vector is expected", name, t.ndim);
    const int64_t R = t.ndim == 2 ? t.shape[0] : 1, Cc = t.ndim == 2 ? t.shape[1] : t.shape[0];
    if (rows < 0) row0 = 0, rows = R;
    if (cols < 0) col0 = 0, cols = Cc;
    if (row0 < 0 || col0 < 0 || rows <= 0 || cols <= 0 || row0 + rows > R || col0 + cols > Cc)
        return jl_set_error(ctx, JL_ERR_INVALID, "%s: slice out of range", name);
    const uint8_t *src = (const uint8_t *)jl_st_data(st, i);
    const float *qb = nullptr;
    int dt = t.dtype;
    if (dt == JL_Q4 || dt == JL_I8) {
        const int j = jl_st_find(st, (std::string(name) + ".qb").c_str());
        if (j < 0) return jl_set_error(ctx, JL_ERR_INVALID, "%s: quantised tensor without %s.qb", name, name);
        // the scales are read as f32 [R, Cc/32] (SafeTensorSupport.java:264-267): check that the sibling really is that
        const StTensor &q = st->tensors[j];
        if ((Cc % 32) || q.dtype != JL_F32 || q.ndim != 2 || q.shape[0] != R || q.shape[1] != Cc / 32)
            return jl_set_error(ctx, JL_ERR_INVALID, "%s.qb: expected F32 block scales of shape [%lld, %lld]", name, (long long)R, (long long)(Cc / 32));
        qb = (const float *)jl_st_data(st, j);
        if ((col0 % 32) || (cols % 32)) return jl_set_error(ctx, JL_ERR_INVALID, "%s: column shard must be a multiple of 32", name);
    }
    std::vector<float> conv;
    if (dt == JL_ST_F16) {
        tensor_to_f32(t, src, conv);
        src = (const uint8_t *)conv.data();
        dt = JL_F32;
    }
    const double bpe = dt == JL_F32 ? 4 : (dt == JL_BF16 ? 2 : (dt == JL_Q4 ? 0.5 : 1));
    const size_t row_bytes = (size_t)(Cc * bpe), sl_bytes = (size_t)(cols * bpe), c0b = (size_t)(col0 * bpe);
    const bool whole_rows = col0 == 0 && cols == Cc;
    HostSlice hs;
    const uint8_t *dptr = src + (size_t)row0 * row_bytes;
    const float *sptr = qb ? (const float *)((const uint8_t *)qb + (size_t)row0 * (size_t)(Cc / 32) * 4) : nullptr;
    if (!whole_rows) {
        hs.data.resize((size_t)rows * sl_bytes);
        for (int64_t r = 0; r < rows; r++) memcpy(hs.data.data() + (size_t)r * sl_bytes, src + (size_t)(row0 + r) * row_bytes + c0b, sl_bytes);
        dptr = hs.data.data();
        if (qb) {
            hs.scales.resize((size_t)rows * (cols / 32));
            for (int64_t r = 0; r < rows; r++)
                memcpy(hs.scales.data() + (size_t)r * (cols / 32), (const uint8_t *)qb + ((size_t)(row0 + r) * (size_t)(Cc / 32) + (size_t)(col0 / 32)) * 4,
                       (size_t)(cols / 32) * 4);
            sptr = hs.scales.data();
        }
    }
    const int64_t id = jl_register_tensor(ctx, dt, rows, cols, dptr, sptr);
    if (id < 0) return JL_ERR_CUDA;
    *id_out = id;
    return JL_OK;
}

extern "C" int jl_model_tp_layout(jl_model *m, jl_dctx *out, int *tp_size); // jl_model.cu
int jl_model_num_experts(jl_model *m);                                          // jl_model.cu

// Registers this rank's shard of every tensor and binds it (the C++ twin of jlama_b200/model.py LlamaModel.__init__).
// ids_out (nullable, capacity ids_cap): the registered tensor ids, for jl_unregister_tensor after jl_model_free.
extern "C" int jl_model_load_safetensors(jl_model *m, jl_ctx *ctx, jl_st *st, int64_t *ids_out, int ids_cap, int *n_ids) {
    if (!m || !ctx || !st) return JL_ERR_INVALID;
    jl_dctx d;
    int layers = 0, tp = 1;
    {
        int rc = jl_model_tp_layout(m, &d, &tp);
        if (rc != JL_OK) return rc;
        layers = d.numberOfLayers;
    }
    int n = 0;
    auto put = [&](int layer, int slot, const std::string &name, int64_t r0, int64_t rows, int64_t c0, int64_t cols, bool optional) -> int {
        if (optional && jl_st_find(st, name.c_str()) < 0) return JL_OK;
        int64_t id = -1;
        int rc = register_slice(ctx, st, name.c_str(), r0, rows, c0, cols, &id);
        if (rc != JL_OK) return rc;
        if (ids_out && n < ids_cap) ids_out[n] = id;
        n++;
        return jl_model_set_tensor(m, layer, slot, id);
    };
    auto put_expert = [&](int layer, int expert, int which, const std::string &name) -> int {
        int64_t id = -1;
        int rc = register_slice(ctx, st, name.c_str(), 0, -1, 0, -1, &id);
        if (rc != JL_OK) return rc;
        if (ids_out && n < ids_cap) ids_out[n] = id;
        n++;
        return jl_model_set_expert_tensor(m, layer, expert, which, id);
    };
    const bool sh = tp > 1;
    const int rank = sh && d.attentionSegmentLength > 0 ? d.attentionSegmentStart / d.attentionSegmentLength : 0;
    const int n_experts = jl_model_num_experts(m);
    int rc = put(-1, JL_T_EMBED, "model.embed_tokens.weight", 0, -1, 0, -1, false);
    if (rc == JL_OK) rc = put(-1, JL_T_OUT_NORM, "model.norm.weight", 0, -1, 0, -1, false);
    if (rc == JL_OK) rc = put(-1, JL_T_LM_HEAD, "lm_head.weight", 0, -1, 0, -1, true); // tied embeddings: absent (LlamaModel.java:152-156)
    for (int L = 0; L < layers && rc == JL_OK; L++) {
        const std::string b = "model.layers." + std::to_string(L) + ".";
        const int64_t ar0 = sh ? d.attentionSegmentStart : 0, ar = sh ? d.attentionSegmentLength : -1;
        const int64_t kr0 = sh ? d.kvSegmentStart : 0, kr = sh ? d.kvSegmentLength : -1;
        const int64_t hr0 = sh ? d.hiddenSegmentStart : 0, hr = sh ? d.hiddenSegmentLength : -1;
        rc = put(L, JL_L_ATTN_NORM, b + "input_layernorm.weight", 0, -1, 0, -1, false);
        if (rc == JL_OK) rc = put(L, JL_L_Q, b + "self_attn.q_proj.weight", ar0, ar, 0, -1, false);
        if (rc == JL_OK) rc = put(L, JL_L_K, b + "self_attn.k_proj.weight", kr0, kr, 0, -1, false);
        if (rc == JL_OK) rc = put(L, JL_L_V, b + "self_attn.v_proj.weight", kr0, kr, 0, -1, false);
        if (rc == JL_OK) rc = put(L, JL_L_O, b + "self_attn.o_proj.weight", 0, -1, ar0, ar, false);
        if (rc == JL_OK) rc = put(L, JL_L_FFN_NORM, b + "post_attention_layernorm.weight", 0, -1, 0, -1, false);
        if (rc == JL_OK && n_experts > 0) {
            // MixtralModel.java:88-105: router + w1 (gate) / w2 (down) / w3 (up) per expert.  Expert parallelism as in model.py:
            // expert e lives whole on rank e % tp, the router is replicated.
            rc = put_expert(L, -1, 0, b + "block_sparse_moe.gate.weight");
            for (int e = 0; e < n_experts && rc == JL_OK; e++) {
                if (e % tp != rank) continue;
                const std::string eb = b + "block_sparse_moe.experts." + std::to_string(e) + ".";
                rc = put_expert(L, e, 0, eb + "w1.weight");
                if (rc == JL_OK) rc = put_expert(L, e, 1, eb + "w2.weight");
                if (rc == JL_OK) rc = put_expert(L, e, 2, eb + "w3.weight");
            }
            continue;
        }
        if (rc == JL_OK) rc = put(L, JL_L_GATE, b + "mlp.gate_proj.weight", hr0, hr, 0, -1, false);
        if (rc == JL_OK) rc = put(L, JL_L_DOWN, b + "mlp.down_proj.weight", 0, -1, hr0, hr, false);
        if (rc == JL_OK) rc = put(L, JL_L_UP, b + "mlp.up_proj.weight", hr0, hr, 0, -1, false);
    }
    if (n_ids) *n_ids = n;
    return rc;
}
