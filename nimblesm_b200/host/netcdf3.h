// nimblesm_b200/host/netcdf3.h — minimal NetCDF-3 (classic CDF-1 and 64-bit-offset CDF-2) reader / writer.
//
// The reference reaches its Genesis / Exodus files through libexodus on libnetcdf (ex_open / ex_get_* /
// ex_put_*, src/nimble_genesis_mesh.cc:62-316, src/nimble_exodus_output.cc:78-396).  Neither library exists
// on the B200 boxes, and every mesh / result file the reference ships for this path is plain NetCDF-3
// 64-bit offset, so this file implements that container format directly:
// header (dimensions, global attributes, variables with attributes), fixed-size variables, and record
// variables along the one unlimited dimension.  Big-endian on disk as the format requires.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

namespace nimble_b200 {
namespace nc3 {

enum Type { BYTE = 1, CHAR = 2, SHORT = 3, INT = 4, FLOAT = 5, DOUBLE = 6 };

int
type_size(int t);

struct Attribute
{
  std::string       name;
  int               type = CHAR;
  std::vector<char> raw;  // host-endian payload, count * type_size bytes
  int64_t
  count() const
  {
    return (int64_t)raw.size() / type_size(type);
  }
  std::string
  as_string() const
  {
    return std::string(raw.begin(), raw.end());
  }
};

struct Variable
{
  std::string            name;
  int                    type = DOUBLE;
  std::vector<int>       dim_ids;
  std::vector<Attribute> attributes;
  bool                   is_record = false;
  int64_t                vsize     = 0;  // bytes of one record slab (record var) or of the whole variable, padded to 4
  int64_t                begin     = 0;  // file offset
  const Attribute*
  find_attribute(const std::string& n) const;
};

// Read-only view of one file, loaded header first; variable payloads are read on request.
class Reader
{
 public:
  explicit Reader(const std::string& path);  // throws std::runtime_error
  ~Reader();
  Reader(const Reader&) = delete;
  Reader&
  operator=(const Reader&) = delete;

  bool
  has_dim(const std::string& n) const
  {
    return dim_index_.count(n) != 0;
  }
  int64_t
  dim(const std::string& n) const;  // length; the record dimension reports the number of records
  int64_t
  dim_or(const std::string& n, int64_t fallback) const
  {
    return has_dim(n) ? dim(n) : fallback;
  }
  bool
  has_var(const std::string& n) const
  {
    return var_index_.count(n) != 0;
  }
  const Variable&
  var(const std::string& n) const;
  std::vector<int64_t>
  shape(const std::string& n) const;
  const Attribute*
  global_attribute(const std::string& n) const;

  // Whole variable (all records for a record variable), converted to the requested C type.
  std::vector<double>
  read_double(const std::string& n) const;
  std::vector<int>
  read_int(const std::string& n) const;
  // CHAR [rows][len] variable -> one trimmed string per row
  std::vector<std::string>
  read_strings(const std::string& n) const;

  int
  version() const
  {
    return version_;
  }
  int64_t
  num_records() const
  {
    return numrecs_;
  }

 private:
  std::vector<char>
  read_raw(const Variable& v, int64_t* n_items) const;
  FILE*                       f_ = nullptr;
  int                         version_ = 1;
  int64_t                     numrecs_ = 0, recsize_ = 0;
  std::vector<std::string>    dim_names_;
  std::vector<int64_t>        dim_lens_;
  int                         rec_dim_ = -1;
  std::map<std::string, int>  dim_index_, var_index_;
  std::vector<Attribute>      gatts_;
  std::vector<Variable>       vars_;
};

// Writer: define everything, end_define() writes the header and zero-fills the fixed-size variables, then
// put_* fill variables and records.  Always CDF-2 (64-bit offset) like the files the reference writes.
class Writer
{
 public:
  explicit Writer(const std::string& path);  // throws std::runtime_error
  ~Writer();
  Writer(const Writer&) = delete;
  Writer&
  operator=(const Writer&) = delete;

  int
  def_dim(const std::string& name, int64_t len);  // len 0 = the unlimited (record) dimension
  int
  dim_id(const std::string& name) const;
  void
  put_global_text(const std::string& name, const std::string& value);
  void
  put_global_int(const std::string& name, int value);
  void
  put_global_float(const std::string& name, float value);
  int
  def_var(const std::string& name, int type, const std::vector<std::string>& dims);
  void
  put_var_text_attribute(const std::string& var, const std::string& name, const std::string& value);
  void
  end_define();

  void
  put_double(const std::string& var, const double* data, int64_t n);
  void
  put_int(const std::string& var, const int* data, int64_t n);
  // CHAR [rows][len]: rows of NUL-padded strings
  void
  put_strings(const std::string& var, const std::vector<std::string>& rows);
  // one record (0-based) of a record variable
  void
  put_record_double(const std::string& var, int64_t record, const double* data, int64_t n);
  void
  flush();
  void
  close();

 private:
  Variable&
  find(const std::string& var);
  int64_t
  fixed_items(const Variable& v) const;
  void
  write_at(int64_t off, const void* p, size_t n);
  void
  ensure_records(int64_t n);
  FILE*                      f_ = nullptr;
  bool                       defining_ = true;
  std::vector<std::string>   dim_names_;
  std::vector<int64_t>       dim_lens_;
  int                        rec_dim_ = -1;
  std::vector<Attribute>     gatts_;
  std::vector<Variable>      vars_;
  std::map<std::string, int> var_index_;
  int64_t                    numrecs_ = 0, recsize_ = 0, rec_begin_ = 0;
};

}  // namespace nc3
}  // namespace nimble_b200
