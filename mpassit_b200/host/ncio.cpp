// ncio.cpp -- NetCDF classic container (CDF-1/2/5), see ncio.hpp.  Written from the published file-format
// grammar ("NetCDF Classic Format Specification": header = magic numrecs dim_list gatt_list var_list).
#include "ncio.hpp"

#include "par.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <thread>

namespace ncio {

namespace {
constexpr uint32_t kTagDim = 0x0A, kTagVar = 0x0B, kTagAtt = 0x0C;

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
inline uint64_t pad4(uint64_t n) { return (n + 3) & ~(uint64_t)3; }

// bounds-checked cursor over the mapped header
struct Cur {
    const uint8_t *p, *end;
    bool ok = true;
    bool need(size_t n) {
        if (!ok || (size_t)(end - p) < n) ok = false;
        return ok;
    }
    uint32_t u32() {
        if (!need(4)) return 0;
        uint32_t v = be32(p);
        p += 4;
        return v;
    }
    uint64_t u64() {
        if (!need(8)) return 0;
        uint64_t v = be64(p);
        p += 8;
        return v;
    }
    uint64_t nonneg(int version) { return version == 5 ? u64() : u32(); }
    std::string name(int version) {
        uint64_t n = nonneg(version);
        if (n > (1u << 20) || !need(pad4(n))) {
            ok = false;
            return "";
        }
        std::string s((const char *)p, n);
        p += pad4(n);
        return s;
    }
};

bool read_atts(Cur &c, int version, std::vector<Att> &out) {
    uint32_t tag = c.u32();
    uint64_t n = c.nonneg(version);
    if (!c.ok) return false;
    if (tag == 0 && n == 0) return true;
    if (tag != kTagAtt) return false;
    for (uint64_t i = 0; i < n && c.ok; ++i) {
        Att a;
        a.name = c.name(version);
        a.type = (int)c.u32();
        a.nelems = c.nonneg(version);
        const size_t es = type_size(a.type);
        if (es == 0) return false;
        const uint64_t bytes = a.nelems * es;
        if (!c.need(pad4(bytes))) return false;
        a.raw.assign(c.p, c.p + bytes);
        c.p += pad4(bytes);
        out.push_back(std::move(a));
    }
    return c.ok;
}

// header serialisation
struct Out {
    std::vector<uint8_t> b;
    int version;
    void u32(uint32_t v) {
        for (int s = 24; s >= 0; s -= 8) b.push_back((uint8_t)(v >> s));
    }
    void u64(uint64_t v) {
        u32((uint32_t)(v >> 32));
        u32((uint32_t)v);
    }
    void nonneg(uint64_t v) { version == 5 ? u64(v) : u32((uint32_t)v); }
    void bytes(const void *p, size_t n) {
        const uint8_t *q = (const uint8_t *)p;
        b.insert(b.end(), q, q + n);
        while (b.size() & 3) b.push_back(0);
    }
    void name(const std::string &s) {
        nonneg(s.size());
        bytes(s.data(), s.size());
    }
    void atts(const std::vector<Att> &a) {
        if (a.empty()) {
            u32(0);
            nonneg(0);
            return;
        }
        u32(kTagAtt);
        nonneg(a.size());
        for (const Att &t : a) {
            name(t.name);
            u32((uint32_t)t.type);
            nonneg(t.nelems);
            bytes(t.raw.data(), t.raw.size());
        }
    }
};

bool pwrite_all(int fd, const void *p, size_t n, uint64_t off, std::string &err) {
    const uint8_t *q = (const uint8_t *)p;
    while (n > 0) {
        ssize_t w = ::pwrite(fd, q, n, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            err = std::string("pwrite: ") + std::strerror(errno);
            return false;
        }
        q += w;
        off += (uint64_t)w;
        n -= (size_t)w;
    }
    return true;
}
}  // namespace

size_t type_size(int t) {
    switch (t) {
        case NC_BYTE: case NC_CHAR: case NC_UBYTE: return 1;
        case NC_SHORT: case NC_USHORT: return 2;
        case NC_INT: case NC_FLOAT: case NC_UINT: return 4;
        case NC_DOUBLE: case NC_INT64: case NC_UINT64: return 8;
        default: return 0;
    }
}

void to_big_endian(void *p, size_t elem, size_t n) {
    if (elem == 4 || elem == 8) {  // the coordinate tables of a 3-km target are 10^7 words: word swaps, a few threads
        par::range((int64_t)n, 1 << 16, [&](int64_t b, int64_t e) {
            if (elem == 4) {
                uint32_t *q = (uint32_t *)p;
                for (int64_t i = b; i < e; ++i) q[i] = __builtin_bswap32(q[i]);
            } else {
                uint64_t *q = (uint64_t *)p;
                for (int64_t i = b; i < e; ++i) q[i] = __builtin_bswap64(q[i]);
            }
        });
        return;
    }
    uint8_t *b = (uint8_t *)p;
    for (size_t i = 0; i < n; ++i, b += elem)
        for (size_t k = 0; k < elem / 2; ++k) std::swap(b[k], b[elem - 1 - k]);
}

double be_number(const uint8_t *p, int type) {
    switch (type) {
        case NC_BYTE: return (double)(int8_t)p[0];
        case NC_CHAR: case NC_UBYTE: return (double)p[0];
        case NC_SHORT: return (double)(int16_t)(((uint16_t)p[0] << 8) | p[1]);
        case NC_USHORT: return (double)(uint16_t)(((uint16_t)p[0] << 8) | p[1]);
        case NC_INT: return (double)(int32_t)be32(p);
        case NC_UINT: return (double)be32(p);
        case NC_FLOAT: { uint32_t u = be32(p); float f; std::memcpy(&f, &u, 4); return (double)f; }
        case NC_DOUBLE: { uint64_t u = be64(p); double d; std::memcpy(&d, &u, 8); return d; }
        case NC_INT64: return (double)(int64_t)be64(p);
        case NC_UINT64: return (double)be64(p);
        default: return 0.0;
    }
}

std::string Att::text() const {
    std::string s((const char *)raw.data(), raw.size());
    while (!s.empty() && s.back() == '\0') s.pop_back();
    return s;
}
double Att::number(uint64_t i) const {
    const size_t es = type_size(type);
    if (es == 0 || (i + 1) * es > raw.size()) return 0.0;
    return be_number(raw.data() + i * es, type);
}
const Att *Var::att(const std::string &n) const {
    for (const Att &a : atts)
        if (a.name == n) return &a;
    return nullptr;
}

// ---------------------------------------------------------------------------
// Reader
// ---------------------------------------------------------------------------
Reader::~Reader() { close(); }
void Reader::close() {
    if (map_) ::munmap((void *)map_, len_);
    if (fd_ >= 0) ::close(fd_);
    map_ = nullptr;
    fd_ = -1;
    len_ = 0;
}

bool Reader::open(const std::string &path, std::string &err) {
    close();
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) {
        err = path + ": " + std::strerror(errno);
        return false;
    }
    struct stat st;
    if (::fstat(fd_, &st) != 0 || st.st_size < 8) {
        err = path + ": not a NetCDF file (too short)";
        return false;
    }
    len_ = (size_t)st.st_size;
    void *m = ::mmap(nullptr, len_, PROT_READ, MAP_SHARED, fd_, 0);
    if (m == MAP_FAILED) {
        err = path + ": mmap: " + std::strerror(errno);
        len_ = 0;
        return false;
    }
    map_ = (const uint8_t *)m;
    if (len_ >= 8 && !std::memcmp(map_, "\x89HDF\r\n\x1a\n", 8)) {
        err = path + ": NetCDF-4 / HDF5 container; this build reads the classic formats only (convert with `nccopy -k cdf5`)";
        return false;
    }
    if (std::memcmp(map_, "CDF", 3) != 0 || (map_[3] != 1 && map_[3] != 2 && map_[3] != 5)) {
        err = path + ": not a NetCDF classic file (magic)";
        return false;
    }
    version = map_[3];
    Cur c{map_ + 4, map_ + len_};
    numrecs = version == 5 ? c.u64() : c.u32();
    const bool streaming = version == 5 ? numrecs == ~(uint64_t)0 : numrecs == 0xFFFFFFFFu;
    // dim_list
    {
        uint32_t tag = c.u32();
        uint64_t n = c.nonneg(version);
        if (!c.ok || !((tag == 0 && n == 0) || tag == kTagDim)) {
            err = path + ": bad dimension list";
            return false;
        }
        for (uint64_t i = 0; i < n && c.ok; ++i) {
            Dim d;
            d.name = c.name(version);
            d.len = c.nonneg(version);
            dims.push_back(std::move(d));
        }
    }
    if (!c.ok || !read_atts(c, version, gatts)) {
        err = path + ": bad global attribute list";
        return false;
    }
    {
        uint32_t tag = c.u32();
        uint64_t n = c.nonneg(version);
        if (!c.ok || !((tag == 0 && n == 0) || tag == kTagVar)) {
            err = path + ": bad variable list";
            return false;
        }
        for (uint64_t i = 0; i < n && c.ok; ++i) {
            Var v;
            v.name = c.name(version);
            uint64_t nd = c.nonneg(version);
            if (nd > 1024) c.ok = false;
            for (uint64_t k = 0; k < nd && c.ok; ++k) {
                uint64_t id = c.nonneg(version);
                if (id >= dims.size()) c.ok = false;
                v.dimids.push_back((int)id);
            }
            if (!c.ok || !read_atts(c, version, v.atts)) {
                err = path + ": bad variable header (" + v.name + ")";
                return false;
            }
            v.type = (int)c.u32();
            v.vsize = c.nonneg(version);
            v.begin = version == 1 ? c.u32() : c.u64();
            v.record = !v.dimids.empty() && dims[v.dimids[0]].len == 0;
            if (type_size(v.type) == 0) c.ok = false;
            vars.push_back(std::move(v));
        }
    }
    if (!c.ok) {
        err = path + ": truncated header";
        return false;
    }
    // vsize is unreliable for variables >= 4 GiB in CDF-1/2 (stored as 2^32-1): recompute from the shape
    int nrec = 0;
    for (Var &v : vars) {
        v.vsize = pad4(count(v) * type_size(v.type));
        nrec += v.record;
    }
    recsize = 0;
    for (Var &v : vars)
        if (v.record) {
            // a lone record variable is not padded
            recsize += nrec == 1 ? count(v) * type_size(v.type) : v.vsize;
        }
    if (streaming) {
        uint64_t first = ~(uint64_t)0;
        for (const Var &v : vars)
            if (v.record && v.begin < first) first = v.begin;
        numrecs = (recsize && first < len_) ? (len_ - first) / recsize : 0;
    }
    for (const Var &v : vars) {
        const uint64_t bytes = count(v) * type_size(v.type);
        const uint64_t need = !v.record ? v.begin + bytes : (numrecs ? v.begin + (numrecs - 1) * recsize + bytes : 0);
        if (need > len_) {
            err = path + ": variable " + v.name + " extends past the end of the file";
            return false;
        }
    }
    return true;
}

const Var *Reader::var(const std::string &name) const {
    for (const Var &v : vars)
        if (v.name == name) return &v;
    return nullptr;
}
const Dim *Reader::dim(const std::string &name) const {
    for (const Dim &d : dims)
        if (d.name == name) return &d;
    return nullptr;
}
const Att *Reader::gatt(const std::string &name) const {
    for (const Att &a : gatts)
        if (a.name == name) return &a;
    return nullptr;
}
uint64_t Reader::count(const Var &v) const {
    uint64_t n = 1;
    for (size_t k = v.record ? 1 : 0; k < v.dimids.size(); ++k) n *= dims[v.dimids[k]].len;
    return n;
}
std::vector<uint64_t> Reader::shape(const Var &v) const {
    std::vector<uint64_t> s;
    for (size_t k = v.record ? 1 : 0; k < v.dimids.size(); ++k) s.push_back(dims[v.dimids[k]].len);
    return s;
}
const uint8_t *Reader::data(const Var &v, uint64_t rec) const {
    if (v.record && rec >= numrecs) return nullptr;
    return map_ + v.begin + (v.record ? rec * recsize : 0);
}
bool Reader::read_doubles(const Var &v, uint64_t first, uint64_t n, double *out, std::string &err) const {
    const uint8_t *p = data(v);
    const size_t es = type_size(v.type);
    if (!p || first + n > count(v) || v.type == NC_CHAR) {
        err = "variable " + v.name + ": cannot read " + std::to_string(n) + " numeric values";
        return false;
    }
    p += first * es;
    const int type = v.type;
    // geometry arrays of a 3-km mesh are 10^7 values: convert on a few threads, the common types without the switch
    par::range((int64_t)n, 1 << 16, [&](int64_t b, int64_t e) {
        if (type == NC_DOUBLE) {
            for (int64_t i = b; i < e; ++i) {
                uint64_t u = be64(p + i * 8);
                std::memcpy(&out[i], &u, 8);
            }
        } else if (type == NC_FLOAT) {
            for (int64_t i = b; i < e; ++i) {
                uint32_t u = be32(p + i * 4);
                float f;
                std::memcpy(&f, &u, 4);
                out[i] = (double)f;
            }
        } else {
            for (int64_t i = b; i < e; ++i) out[i] = be_number(p + i * es, type);
        }
    });
    return true;
}
bool Reader::read_ints(const Var &v, uint64_t first, uint64_t n, int32_t *out, std::string &err) const {
    const uint8_t *p = data(v);
    if (!p || first + n > count(v) || v.type != NC_INT) {
        err = "variable " + v.name + ": cannot read " + std::to_string(n) + " int values";
        return false;
    }
    p += first * 4;
    par::range((int64_t)n, 1 << 16, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) out[i] = (int32_t)be32(p + i * 4);
    });
    return true;
}

// ---------------------------------------------------------------------------
// Writer
// ---------------------------------------------------------------------------
Writer::~Writer() {
    if (fd_ >= 0) ::close(fd_);
}
int Writer::def_dim(const std::string &name, uint64_t len) {
    dims_.push_back(Dim{name, len});
    return (int)dims_.size() - 1;
}
Att *Writer::new_att(int varid, const std::string &name) {
    std::vector<Att> &list = varid < 0 ? gatts_ : vars_[varid].atts;
    for (Att &a : list)
        if (a.name == name) return &a;  // nf90_put_att on an existing name overwrites it
    list.push_back(Att{});
    list.back().name = name;
    return &list.back();
}
void Writer::att_raw(int varid, const Att &a) {
    Att *t = new_att(varid, a.name);
    t->type = a.type;
    t->nelems = a.nelems;
    t->raw = a.raw;
}
void Writer::att_text(int varid, const std::string &name, const std::string &value) {
    Att *a = new_att(varid, name);
    a->type = NC_CHAR;
    a->nelems = value.size();
    a->raw.assign(value.begin(), value.end());
}
void Writer::att_int(int varid, const std::string &name, int32_t value) {
    Att *a = new_att(varid, name);
    a->type = NC_INT;
    a->nelems = 1;
    a->raw.resize(4);
    std::memcpy(a->raw.data(), &value, 4);
    to_big_endian(a->raw.data(), 4, 1);
}
void Writer::att_double(int varid, const std::string &name, double value) {
    Att *a = new_att(varid, name);
    a->type = NC_DOUBLE;
    a->nelems = 1;
    a->raw.resize(8);
    std::memcpy(a->raw.data(), &value, 8);
    to_big_endian(a->raw.data(), 8, 1);
}
int Writer::def_var(const std::string &name, int type, const std::vector<int> &dimids) {
    Var v;
    v.name = name;
    v.type = type;
    v.dimids = dimids;
    v.record = !dimids.empty() && dims_[dimids[0]].len == 0;
    vars_.push_back(std::move(v));
    return (int)vars_.size() - 1;
}
uint64_t Writer::var_bytes(int varid) const {
    const Var &v = vars_[varid];
    uint64_t n = type_size(v.type);
    for (size_t k = v.record ? 1 : 0; k < v.dimids.size(); ++k) n *= dims_[v.dimids[k]].len;
    return n;
}

bool Writer::enddef(const std::string &path, uint64_t numrecs, bool create_file, std::string &err) {
    int nrec = 0;
    uint64_t biggest = 0;
    for (size_t i = 0; i < vars_.size(); ++i) {
        vars_[i].vsize = pad4(var_bytes((int)i));
        nrec += vars_[i].record;
        if (vars_[i].vsize > biggest) biggest = vars_[i].vsize;
    }
    // CDF-2 holds variables (records) below 4 GiB each; beyond that the 64-bit-data format is needed
    version_ = want_version_ ? want_version_ : (biggest >= ((uint64_t)1 << 32) - 4 ? 5 : 2);
    if (version_ != 1 && version_ != 2 && version_ != 5) {
        err = "unsupported NetCDF classic version " + std::to_string(version_);
        return false;
    }
    auto serialise = [&](uint64_t nrecs) {
        Out o;
        o.version = version_;
        o.b = {'C', 'D', 'F', (uint8_t)version_};
        o.nonneg(nrecs);
        if (dims_.empty()) {
            o.u32(0);
            o.nonneg(0);
        } else {
            o.u32(kTagDim);
            o.nonneg(dims_.size());
            for (const Dim &d : dims_) {
                o.name(d.name);
                o.nonneg(d.len);
            }
        }
        o.atts(gatts_);
        if (vars_.empty()) {
            o.u32(0);
            o.nonneg(0);
        } else {
            o.u32(kTagVar);
            o.nonneg(vars_.size());
            for (const Var &v : vars_) {
                o.name(v.name);
                o.nonneg(v.dimids.size());
                for (int d : v.dimids) o.nonneg((uint64_t)d);
                o.atts(v.atts);
                o.u32((uint32_t)v.type);
                o.nonneg(version_ == 5 ? v.vsize : (v.vsize > 0xFFFFFFFFull ? 0xFFFFFFFFull : v.vsize));
                if (version_ == 1) o.u32((uint32_t)v.begin);
                else o.u64(v.begin);
            }
        }
        return o.b;
    };
    // the header's length does not depend on the offsets it stores: lay out, then serialise again
    const uint64_t hdr = pad4(serialise(numrecs).size());
    uint64_t off = hdr;
    for (Var &v : vars_)
        if (!v.record) {
            v.begin = off;
            off += v.vsize;
        }
    uint64_t recsize = 0;
    for (size_t i = 0; i < vars_.size(); ++i)
        if (vars_[i].record) {
            vars_[i].begin = off + recsize;
            recsize += nrec == 1 ? var_bytes((int)i) : vars_[i].vsize;
        }
    file_size_ = off + recsize * numrecs;
    if (version_ == 1 && file_size_ > 0x7FFFFFFFull) {
        err = "file too large for CDF-1";
        return false;
    }
    recsize_ = recsize;
    const std::vector<uint8_t> h = serialise(numrecs);
    fd_ = ::open(path.c_str(), create_file ? (O_RDWR | O_CREAT | O_TRUNC) : O_RDWR, 0644);
    if (fd_ < 0) {
        err = path + ": " + std::strerror(errno);
        return false;
    }
    if (create_file) {
        if (::ftruncate(fd_, (off_t)file_size_) != 0) {
            err = path + ": ftruncate: " + std::strerror(errno);
            return false;
        }
        if (!pwrite_all(fd_, h.data(), h.size(), 0, err)) return false;
    }
    return true;
}

uint64_t Writer::var_offset(int varid, uint64_t rec) const {
    const Var &v = vars_[varid];
    return v.begin + (v.record ? rec * recsize_ : 0);
}
bool Writer::write_raw(int varid, uint64_t rec, uint64_t byte_off, const void *p, size_t n, std::string &err) {
    if (fd_ < 0) {
        err = "write before enddef";
        return false;
    }
    if (byte_off + n > var_bytes(varid)) {
        err = "write past the end of variable " + vars_[varid].name;
        return false;
    }
    const uint64_t off = var_offset(varid, rec) + byte_off;
    // Large blocks (a rank's slab of a 3-D field is 10^8 bytes) are cut across a few threads: one pwrite copies
    // into the page cache at a single core's memcpy speed, and the regions are disjoint by construction.
    constexpr size_t kParallelFrom = (size_t)8 << 20, kAlign = (size_t)1 << 20;
    unsigned nt = std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
    if (n < kParallelFrom || nt < 2) return pwrite_all(fd_, p, n, off, err);
    const size_t chunk = ((n + nt - 1) / nt + kAlign - 1) & ~(kAlign - 1);
    nt = (unsigned)((n + chunk - 1) / chunk);
    std::vector<std::string> errs(nt);
    std::vector<char> ok(nt, 1);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            const size_t a = (size_t)t * chunk, len = std::min(chunk, n - a);
            ok[t] = pwrite_all(fd_, (const uint8_t *)p + a, len, off + a, errs[t]) ? 1 : 0;
        });
    for (auto &x : th) x.join();
    for (unsigned t = 0; t < nt; ++t)
        if (!ok[t]) {
            err = errs[t];
            return false;
        }
    return true;
}
bool Writer::put_doubles(int varid, const double *v, uint64_t n, std::string &err) {
    const Var &var = vars_[varid];
    if (var.type == NC_FLOAT) {
        std::vector<float> f(n);
        par::range((int64_t)n, 1 << 16, [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; ++i) f[i] = (float)v[i];
        });
        to_big_endian(f.data(), 4, n);
        return write_raw(varid, 0, 0, f.data(), n * 4, err);
    }
    if (var.type == NC_DOUBLE) {
        std::vector<double> d(v, v + n);
        to_big_endian(d.data(), 8, n);
        return write_raw(varid, 0, 0, d.data(), n * 8, err);
    }
    err = "put_doubles: variable " + var.name + " is not float/double";
    return false;
}
bool Writer::put_ints(int varid, const int32_t *v, uint64_t n, std::string &err) {
    if (vars_[varid].type != NC_INT) {
        err = "put_ints: variable " + vars_[varid].name + " is not int";
        return false;
    }
    std::vector<int32_t> d(v, v + n);
    to_big_endian(d.data(), 4, n);
    return write_raw(varid, 0, 0, d.data(), n * 4, err);
}
bool Writer::put_text(int varid, const std::string &s, std::string &err) {
    if (vars_[varid].type != NC_CHAR) {
        err = "put_text: variable " + vars_[varid].name + " is not char";
        return false;
    }
    std::string t = s;
    t.resize(var_bytes(varid), '\0');
    return write_raw(varid, 0, 0, t.data(), t.size(), err);
}
bool Writer::close(std::string &err) {
    if (fd_ >= 0 && ::close(fd_) != 0) {
        err = std::string("close: ") + std::strerror(errno);
        fd_ = -1;
        return false;
    }
    fd_ = -1;
    return true;
}

}  // namespace ncio

// ---------------------------------------------------------------------------
// C test hooks (include/mpassit_host.h)
// ---------------------------------------------------------------------------
#include <cstdio>

#include "../../include/mpassit_host.h"

namespace {
void put_err(char *err, size_t n, const std::string &m) {
    if (err && n) std::snprintf(err, n, "%s", m.c_str());
}
}  // namespace

extern "C" {

int mpassit_nc_describe(const char *path, char *out, size_t outlen, char *err, size_t errlen) {
    ncio::Reader r;
    std::string why;
    if (!path || !r.open(path, why)) {
        put_err(err, errlen, why);
        return 1;
    }
    std::string s = "version " + std::to_string(r.version) + " numrecs " + std::to_string(r.numrecs) + "\n";
    for (const ncio::Dim &d : r.dims) s += "dim " + d.name + " " + std::to_string(d.len) + "\n";
    for (const ncio::Att &a : r.gatts) s += "gatt " + a.name + " " + std::to_string(a.type) + " " + std::to_string(a.nelems) + "\n";
    for (const ncio::Var &v : r.vars) {
        s += "var " + v.name + " " + std::to_string(v.type) + " " + std::to_string(v.begin);
        for (int d : v.dimids) s += " " + r.dims[d].name;
        s += "\n";
        for (const ncio::Att &a : v.atts)
            s += "vatt " + v.name + " " + a.name + " " + std::to_string(a.type) + " " + std::to_string(a.nelems) + "\n";
    }
    if (out && outlen) std::snprintf(out, outlen, "%s", s.c_str());
    return 0;
}

int mpassit_nc_get(const char *path, const char *var, int64_t rec, int64_t first, int64_t n, double *out, char *err,
                   size_t errlen) {
    ncio::Reader r;
    std::string why;
    if (!path || !var || !r.open(path, why)) {
        put_err(err, errlen, why);
        return 1;
    }
    const ncio::Var *v = r.var(var);
    if (!v) {
        put_err(err, errlen, std::string("NetCDF: Variable not found: ") + var);
        return 2;
    }
    const uint8_t *p = r.data(*v, (uint64_t)rec);
    if (!p || first < 0 || n < 0 || (uint64_t)(first + n) > r.count(*v)) {
        put_err(err, errlen, "NetCDF: Start+count exceeds dimension bound");
        return 3;
    }
    const size_t es = ncio::type_size(v->type);
    for (int64_t i = 0; i < n; ++i) out[i] = ncio::be_number(p + (first + i) * es, v->type);
    return 0;
}

int mpassit_nc_copy(const char *src, const char *dst, int version, char *err, size_t errlen) {
    ncio::Reader r;
    std::string why;
    if (!src || !dst || !r.open(src, why)) {
        put_err(err, errlen, why);
        return 1;
    }
    ncio::Writer w(version);
    for (const ncio::Dim &d : r.dims) w.def_dim(d.name, d.len);
    for (const ncio::Att &a : r.gatts) w.att_raw(-1, a);
    for (const ncio::Var &v : r.vars) {
        const int id = w.def_var(v.name, v.type, v.dimids);
        for (const ncio::Att &a : v.atts) w.att_raw(id, a);
    }
    if (!w.enddef(dst, r.numrecs, true, why)) {
        put_err(err, errlen, why);
        return 2;
    }
    for (size_t i = 0; i < r.vars.size(); ++i) {
        const ncio::Var &v = r.vars[i];
        const uint64_t bytes = r.count(v) * ncio::type_size(v.type);
        const uint64_t nrec = v.record ? r.numrecs : 1;
        for (uint64_t k = 0; k < nrec; ++k)
            if (!w.write_raw((int)i, k, 0, r.data(v, k), bytes, why)) {
                put_err(err, errlen, why);
                return 3;
            }
    }
    if (!w.close(why)) {
        put_err(err, errlen, why);
        return 4;
    }
    return 0;
}

}  // extern "C"
