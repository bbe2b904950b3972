// ncio.hpp -- the NetCDF "classic" container (CDF-1, CDF-2 / 64-bit offset, CDF-5 / 64-bit data) read and
// written directly: this image has no NetCDF / HDF5 library, and the engine wants file bytes, not copies.
//
// Reader: the file is mapped read-only; a variable's record is a contiguous big-endian block inside the
//   mapping, handed to the engine as it lies (MPAS writes `(Time, nCells, nVertLevels)`, i.e. [cell][level]
//   level-fastest per record -- the layout mprg_apply consumes; the transpose of input_data.F90:653-655 and
//   the per-PET whole-variable read of :645 have no counterpart here).
// Writer: the header is laid out once; every variable of the reference's output file has `Time` as its slowest
//   dimension (write_data.F90:281-380), so with one record each variable is one contiguous [lev][j][i] block and
//   a rank's row slab is one run per level -- written with pwrite at its own offset, no gather to PET 0
//   (write_data.F90:996-1496 funnels every field through ESMF_FieldGather + nf90_put_var on PET 0).
//
// NetCDF-4 / HDF5 files (what nf90_create(NF90_NETCDF4) at write_data.F90:170 produces, and what MPAS writes
// with io_type="netcdf4") are NOT handled: same data model, different container; convert with `nccopy -k cdf5`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ncio {

enum Type { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6,
            NC_UBYTE = 7, NC_USHORT = 8, NC_UINT = 9, NC_INT64 = 10, NC_UINT64 = 11 };
size_t type_size(int t);

struct Att {
    std::string name;
    int type = NC_CHAR;
    uint64_t nelems = 0;
    std::vector<uint8_t> raw;  // big-endian values, unpadded
    std::string text() const;  // NC_CHAR payload, trailing NULs dropped
    double number(uint64_t i = 0) const;
};
struct Dim {
    std::string name;
    uint64_t len = 0;  // 0 = the record (unlimited) dimension
};
struct Var {
    std::string name;
    std::vector<int> dimids;
    std::vector<Att> atts;
    int type = NC_FLOAT;
    uint64_t vsize = 0, begin = 0;
    bool record = false;
    const Att *att(const std::string &n) const;
};

class Reader {
  public:
    Reader() = default;
    ~Reader();
    Reader(const Reader &) = delete;
    Reader &operator=(const Reader &) = delete;
    bool open(const std::string &path, std::string &err);
    void close();
    int version = 0;
    uint64_t numrecs = 0, recsize = 0;
    std::vector<Dim> dims;
    std::vector<Att> gatts;
    std::vector<Var> vars;
    const Var *var(const std::string &name) const;
    const Dim *dim(const std::string &name) const;
    const Att *gatt(const std::string &name) const;
    // elements of one record of `v` (the whole variable when it has no record dimension)
    uint64_t count(const Var &v) const;
    // extents without the record dimension, slowest first
    std::vector<uint64_t> shape(const Var &v) const;
    // big-endian bytes of record `rec` inside the mapping
    const uint8_t *data(const Var &v, uint64_t rec = 0) const;
    // elements [first, first+n) of record 0 converted to native doubles / int32 (geometry, small arrays)
    bool read_doubles(const Var &v, uint64_t first, uint64_t n, double *out, std::string &err) const;
    bool read_ints(const Var &v, uint64_t first, uint64_t n, int32_t *out, std::string &err) const;
    const uint8_t *base() const { return map_; }
    size_t size() const { return len_; }

  private:
    const uint8_t *map_ = nullptr;
    size_t len_ = 0;
    int fd_ = -1;
};

class Writer {
  public:
    // version 2 (64-bit offsets) or 5 (64-bit data); 0 = pick 5 only when a variable needs it
    explicit Writer(int version = 0) : want_version_(version) {}
    ~Writer();
    int def_dim(const std::string &name, uint64_t len);  // len 0 = record dimension
    // varid < 0: global attribute
    void att_raw(int varid, const Att &a);  // big-endian payload as read
    void att_text(int varid, const std::string &name, const std::string &value);
    void att_int(int varid, const std::string &name, int32_t value);
    void att_double(int varid, const std::string &name, double value);
    int def_var(const std::string &name, int type, const std::vector<int> &dimids);
    // lay out the header for `numrecs` records; create_file: write it and size the file (the writing rank),
    // otherwise just open the file the writing rank created (same definitions give the same layout)
    bool enddef(const std::string &path, uint64_t numrecs, bool create_file, std::string &err);
    int version() const { return version_; }
    uint64_t var_offset(int varid, uint64_t rec = 0) const;
    uint64_t var_bytes(int varid) const;  // unpadded bytes of one record
    uint64_t file_size() const { return file_size_; }
    // raw big-endian bytes at `byte_off` inside record `rec` of the variable
    bool write_raw(int varid, uint64_t rec, uint64_t byte_off, const void *p, size_t n, std::string &err);
    // native values converted to the variable's type and byte order (small arrays: coordinates, times)
    bool put_doubles(int varid, const double *v, uint64_t n, std::string &err);
    bool put_ints(int varid, const int32_t *v, uint64_t n, std::string &err);
    bool put_text(int varid, const std::string &s, std::string &err);
    bool close(std::string &err);

  private:
    std::vector<Dim> dims_;
    std::vector<Att> gatts_;
    std::vector<Var> vars_;
    int want_version_ = 0, version_ = 2, fd_ = -1;
    uint64_t file_size_ = 0, recsize_ = 0;
    Att *new_att(int varid, const std::string &name);
};

// one big-endian value of NetCDF type `type` as a double
double be_number(const uint8_t *p, int type);
// host byte swap of small arrays
void to_big_endian(void *p, size_t elem, size_t n);

}  // namespace ncio
