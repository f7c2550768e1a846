// nimblesm_b200/host/netcdf3.cc — see netcdf3.h.  File format: "The NetCDF Classic Format Specification"
// (header grammar: magic numrecs dim_list gatt_list var_list; all integers big-endian; names and values
// padded to 4 bytes; record variables interleaved record by record).
#include "netcdf3.h"

#include <cstring>
#include <stdexcept>

namespace nimble_b200 {
namespace nc3 {

namespace {

constexpr int32_t kTagDim = 0x0A, kTagVar = 0x0B, kTagAtt = 0x0C;

int64_t
pad4(int64_t n)
{
  return (n + 3) & ~(int64_t)3;
}

void
swap_bytes(char* p, int size, int64_t count)
{
  if (size == 1) return;
  for (int64_t i = 0; i < count; ++i) {
    char* q = p + i * size;
    for (int a = 0, b = size - 1; a < b; ++a, --b) {
      char t = q[a];
      q[a]   = q[b];
      q[b]   = t;
    }
  }
}

struct In
{
  FILE* f;
  void
  raw(void* p, size_t n)
  {
    if (n && fread(p, 1, n, f) != n) throw std::runtime_error("netcdf3: unexpected end of file in header");
  }
  int32_t
  i32()
  {
    unsigned char b[4];
    raw(b, 4);
    return (int32_t)(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3]);
  }
  int64_t
  i64()
  {
    uint64_t hi = (uint32_t)i32();
    uint64_t lo = (uint32_t)i32();
    return (int64_t)((hi << 32) | lo);
  }
  std::string
  name()
  {
    int32_t n = i32();
    if (n < 0 || n > (1 << 20)) throw std::runtime_error("netcdf3: bad name length");
    std::string s((size_t)pad4(n), '\0');
    raw(&s[0], s.size());
    s.resize((size_t)n);
    return s;
  }
  std::vector<Attribute>
  attributes()
  {
    std::vector<Attribute> out;
    int32_t                tag = i32(), n = i32();
    if (tag == 0 && n == 0) return out;
    if (tag != kTagAtt) throw std::runtime_error("netcdf3: attribute list expected");
    for (int32_t i = 0; i < n; ++i) {
      Attribute a;
      a.name        = name();
      a.type        = i32();
      int32_t count = i32();
      int     ts    = type_size(a.type);
      a.raw.resize((size_t)pad4((int64_t)count * ts));
      raw(a.raw.data(), a.raw.size());
      a.raw.resize((size_t)count * ts);
      swap_bytes(a.raw.data(), ts, count);
      out.push_back(std::move(a));
    }
    return out;
  }
};

struct Out
{
  std::vector<char> buf;
  void
  raw(const void* p, size_t n)
  {
    const char* c = (const char*)p;
    buf.insert(buf.end(), c, c + n);
  }
  void
  i32(int32_t v)
  {
    unsigned char b[4] = {(unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v};
    raw(b, 4);
  }
  void
  i64(int64_t v)
  {
    i32((int32_t)((uint64_t)v >> 32));
    i32((int32_t)((uint64_t)v & 0xffffffffu));
  }
  void
  name(const std::string& s)
  {
    i32((int32_t)s.size());
    raw(s.data(), s.size());
    static const char zeros[4] = {0, 0, 0, 0};
    raw(zeros, (size_t)(pad4((int64_t)s.size()) - (int64_t)s.size()));
  }
  void
  attributes(const std::vector<Attribute>& atts)
  {
    if (atts.empty()) {
      i32(0);
      i32(0);
      return;
    }
    i32(kTagAtt);
    i32((int32_t)atts.size());
    for (const Attribute& a : atts) {
      name(a.name);
      i32(a.type);
      i32((int32_t)a.count());
      std::vector<char> tmp = a.raw;
      swap_bytes(tmp.data(), type_size(a.type), a.count());
      tmp.resize((size_t)pad4((int64_t)tmp.size()), 0);
      raw(tmp.data(), tmp.size());
    }
  }
};

template <class T>
std::vector<T>
convert(const std::vector<char>& raw, int type, int64_t n)
{
  std::vector<T> out((size_t)n);
  const char*    p = raw.data();
  for (int64_t i = 0; i < n; ++i) {
    switch (type) {
      case BYTE: out[i] = (T)(signed char)p[i]; break;
      case CHAR: out[i] = (T)(unsigned char)p[i]; break;
      case SHORT: {
        int16_t v;
        memcpy(&v, p + 2 * i, 2);
        out[i] = (T)v;
        break;
      }
      case INT: {
        int32_t v;
        memcpy(&v, p + 4 * i, 4);
        out[i] = (T)v;
        break;
      }
      case FLOAT: {
        float v;
        memcpy(&v, p + 4 * i, 4);
        out[i] = (T)v;
        break;
      }
      default: {
        double v;
        memcpy(&v, p + 8 * i, 8);
        out[i] = (T)v;
      }
    }
  }
  return out;
}

}  // namespace

int
type_size(int t)
{
  switch (t) {
    case BYTE:
    case CHAR: return 1;
    case SHORT: return 2;
    case INT:
    case FLOAT: return 4;
    case DOUBLE: return 8;
  }
  throw std::runtime_error("netcdf3: unknown external type");
}

const Attribute*
Variable::find_attribute(const std::string& n) const
{
  for (const Attribute& a : attributes)
    if (a.name == n) return &a;
  return nullptr;
}

// ---- Reader -------------------------------------------------------------------------------------------
Reader::Reader(const std::string& path)
{
  f_ = fopen(path.c_str(), "rb");
  if (!f_) throw std::runtime_error("netcdf3: cannot open " + path);
  In   in{f_};
  char magic[4];
  in.raw(magic, 4);
  if (magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' || (magic[3] != 1 && magic[3] != 2))
    throw std::runtime_error("netcdf3: " + path + " is not a NetCDF-3 classic / 64-bit-offset file");
  version_ = magic[3];
  numrecs_ = (uint32_t)in.i32();
  {
    int32_t tag = in.i32(), n = in.i32();
    if (!(tag == 0 && n == 0)) {
      if (tag != kTagDim) throw std::runtime_error("netcdf3: dimension list expected");
      for (int32_t i = 0; i < n; ++i) {
        dim_names_.push_back(in.name());
        dim_lens_.push_back((uint32_t)in.i32());
        if (dim_lens_.back() == 0) rec_dim_ = i;
        dim_index_[dim_names_.back()] = i;
      }
    }
  }
  gatts_ = in.attributes();
  {
    int32_t tag = in.i32(), n = in.i32();
    if (!(tag == 0 && n == 0)) {
      if (tag != kTagVar) throw std::runtime_error("netcdf3: variable list expected");
      for (int32_t i = 0; i < n; ++i) {
        Variable v;
        v.name     = in.name();
        int32_t nd = in.i32();
        for (int32_t d = 0; d < nd; ++d) v.dim_ids.push_back(in.i32());
        v.attributes = in.attributes();
        v.type       = in.i32();
        v.vsize      = (uint32_t)in.i32();
        v.begin      = version_ == 2 ? in.i64() : (int64_t)(uint32_t)in.i32();
        v.is_record  = !v.dim_ids.empty() && v.dim_ids[0] == rec_dim_;
        // vsize is not trusted for large variables (it saturates at 2^32 - 4): recompute
        int64_t items = 1;
        for (size_t d = v.is_record ? 1 : 0; d < v.dim_ids.size(); ++d) items *= dim_lens_[v.dim_ids[d]];
        v.vsize = pad4(items * type_size(v.type));
        var_index_[v.name] = (int)vars_.size();
        vars_.push_back(std::move(v));
      }
    }
  }
  int n_rec_vars = 0;
  for (const Variable& v : vars_)
    if (v.is_record) {
      recsize_ += v.vsize;
      ++n_rec_vars;
    }
  if (n_rec_vars == 1)  // a lone record variable is not padded
    for (const Variable& v : vars_)
      if (v.is_record) {
        int64_t items = 1;
        for (size_t d = 1; d < v.dim_ids.size(); ++d) items *= dim_lens_[v.dim_ids[d]];
        recsize_ = items * type_size(v.type);
      }
}

Reader::~Reader()
{
  if (f_) fclose(f_);
}

int64_t
Reader::dim(const std::string& n) const
{
  auto it = dim_index_.find(n);
  if (it == dim_index_.end()) throw std::runtime_error("netcdf3: no dimension " + n);
  return it->second == rec_dim_ ? numrecs_ : dim_lens_[it->second];
}

const Variable&
Reader::var(const std::string& n) const
{
  auto it = var_index_.find(n);
  if (it == var_index_.end()) throw std::runtime_error("netcdf3: no variable " + n);
  return vars_[it->second];
}

std::vector<int64_t>
Reader::shape(const std::string& n) const
{
  const Variable&      v = var(n);
  std::vector<int64_t> s;
  for (int d : v.dim_ids) s.push_back(d == rec_dim_ ? numrecs_ : dim_lens_[d]);
  return s;
}

const Attribute*
Reader::global_attribute(const std::string& n) const
{
  for (const Attribute& a : gatts_)
    if (a.name == n) return &a;
  return nullptr;
}

std::vector<char>
Reader::read_raw(const Variable& v, int64_t* n_items) const
{
  const int ts    = type_size(v.type);
  int64_t   items = 1;
  for (size_t d = v.is_record ? 1 : 0; d < v.dim_ids.size(); ++d) items *= dim_lens_[v.dim_ids[d]];
  const int64_t     nrec = v.is_record ? numrecs_ : 1;
  std::vector<char> raw((size_t)(items * nrec * ts));
  for (int64_t r = 0; r < nrec; ++r) {
    if (fseeko(f_, (off_t)(v.begin + r * recsize_), SEEK_SET) != 0) throw std::runtime_error("netcdf3: seek failed");
    const size_t want = (size_t)(items * ts);
    if (want && fread(raw.data() + r * items * ts, 1, want, f_) != want)
      throw std::runtime_error("netcdf3: short read of variable " + v.name);
  }
  swap_bytes(raw.data(), ts, items * nrec);
  *n_items = items * nrec;
  return raw;
}

std::vector<double>
Reader::read_double(const std::string& n) const
{
  const Variable&   v = var(n);
  int64_t           k = 0;
  std::vector<char> raw = read_raw(v, &k);
  return convert<double>(raw, v.type, k);
}

std::vector<int>
Reader::read_int(const std::string& n) const
{
  const Variable&   v = var(n);
  int64_t           k = 0;
  std::vector<char> raw = read_raw(v, &k);
  return convert<int>(raw, v.type, k);
}

std::vector<std::string>
Reader::read_strings(const std::string& n) const
{
  const Variable& v = var(n);
  if (v.type != CHAR) throw std::runtime_error("netcdf3: " + n + " is not a character variable");
  int64_t                  k   = 0;
  std::vector<char>        raw = read_raw(v, &k);
  std::vector<int64_t>     sh  = shape(n);
  const int64_t            len = sh.empty() ? k : sh.back();
  const int64_t            rows = len ? k / len : 0;
  std::vector<std::string> out;
  for (int64_t r = 0; r < rows; ++r) {
    std::string s(raw.data() + r * len, (size_t)len);
    size_t      z = s.find('\0');
    if (z != std::string::npos) s.resize(z);
    while (!s.empty() && (s.back() == ' ')) s.pop_back();
    out.push_back(s);
  }
  return out;
}

// ---- Writer -------------------------------------------------------------------------------------------
Writer::Writer(const std::string& path)
{
  f_ = fopen(path.c_str(), "wb+");
  if (!f_) throw std::runtime_error("netcdf3: cannot create " + path);
}

Writer::~Writer()
{
  close();
}

void
Writer::close()
{
  if (f_) {
    fclose(f_);
    f_ = nullptr;
  }
}

void
Writer::flush()
{
  if (f_) fflush(f_);
}

int
Writer::def_dim(const std::string& name, int64_t len)
{
  if (!defining_) throw std::runtime_error("netcdf3: def_dim after end_define");
  if (len == 0) {
    if (rec_dim_ >= 0) throw std::runtime_error("netcdf3: only one unlimited dimension is allowed");
    rec_dim_ = (int)dim_names_.size();
  }
  dim_names_.push_back(name);
  dim_lens_.push_back(len);
  return (int)dim_names_.size() - 1;
}

int
Writer::dim_id(const std::string& name) const
{
  for (size_t i = 0; i < dim_names_.size(); ++i)
    if (dim_names_[i] == name) return (int)i;
  throw std::runtime_error("netcdf3: undefined dimension " + name);
}

void
Writer::put_global_text(const std::string& name, const std::string& value)
{
  Attribute a;
  a.name = name, a.type = CHAR;
  a.raw.assign(value.begin(), value.end());
  gatts_.push_back(a);
}

void
Writer::put_global_int(const std::string& name, int value)
{
  Attribute a;
  a.name = name, a.type = INT;
  a.raw.resize(4);
  int32_t v = value;
  memcpy(a.raw.data(), &v, 4);
  gatts_.push_back(a);
}

void
Writer::put_global_float(const std::string& name, float value)
{
  Attribute a;
  a.name = name, a.type = FLOAT;
  a.raw.resize(4);
  memcpy(a.raw.data(), &value, 4);
  gatts_.push_back(a);
}

int
Writer::def_var(const std::string& name, int type, const std::vector<std::string>& dims)
{
  if (!defining_) throw std::runtime_error("netcdf3: def_var after end_define");
  Variable v;
  v.name = name, v.type = type;
  for (const std::string& d : dims) v.dim_ids.push_back(dim_id(d));
  v.is_record = !v.dim_ids.empty() && v.dim_ids[0] == rec_dim_;
  var_index_[name] = (int)vars_.size();
  vars_.push_back(v);
  return (int)vars_.size() - 1;
}

void
Writer::put_var_text_attribute(const std::string& var, const std::string& name, const std::string& value)
{
  Attribute a;
  a.name = name, a.type = CHAR;
  a.raw.assign(value.begin(), value.end());
  find(var).attributes.push_back(a);
}

Variable&
Writer::find(const std::string& var)
{
  auto it = var_index_.find(var);
  if (it == var_index_.end()) throw std::runtime_error("netcdf3: undefined variable " + var);
  return vars_[it->second];
}

int64_t
Writer::fixed_items(const Variable& v) const
{
  int64_t items = 1;
  for (size_t d = v.is_record ? 1 : 0; d < v.dim_ids.size(); ++d) items *= dim_lens_[v.dim_ids[d]];
  return items;
}

void
Writer::write_at(int64_t off, const void* p, size_t n)
{
  if (fseeko(f_, (off_t)off, SEEK_SET) != 0 || (n && fwrite(p, 1, n, f_) != n))
    throw std::runtime_error("netcdf3: write failed");
}

void
Writer::end_define()
{
  if (!defining_) return;
  defining_ = false;
  int n_rec_vars = 0;
  for (Variable& v : vars_) {
    v.vsize = pad4(fixed_items(v) * type_size(v.type));
    if (v.is_record) ++n_rec_vars;
  }
  // two passes: header size depends only on names / counts, not on the offsets' values
  auto build = [&](Out& o) {
    const char magic[4] = {'C', 'D', 'F', 2};
    o.raw(magic, 4);
    o.i32((int32_t)numrecs_);
    if (dim_names_.empty()) {
      o.i32(0), o.i32(0);
    } else {
      o.i32(kTagDim);
      o.i32((int32_t)dim_names_.size());
      for (size_t i = 0; i < dim_names_.size(); ++i) {
        o.name(dim_names_[i]);
        o.i32((int32_t)dim_lens_[i]);
      }
    }
    o.attributes(gatts_);
    if (vars_.empty()) {
      o.i32(0), o.i32(0);
    } else {
      o.i32(kTagVar);
      o.i32((int32_t)vars_.size());
      for (const Variable& v : vars_) {
        o.name(v.name);
        o.i32((int32_t)v.dim_ids.size());
        for (int d : v.dim_ids) o.i32(d);
        o.attributes(v.attributes);
        o.i32(v.type);
        o.i32((int32_t)(v.vsize > 0xfffffffcLL ? 0xfffffffcLL : v.vsize));
        o.i64(v.begin);
      }
    }
  };
  Out probe;
  build(probe);
  int64_t off = pad4((int64_t)probe.buf.size());
  for (Variable& v : vars_)
    if (!v.is_record) {
      v.begin = off;
      off += v.vsize;
    }
  rec_begin_ = off;
  recsize_   = 0;
  for (Variable& v : vars_)
    if (v.is_record) {
      v.begin = off;
      off += v.vsize;
      recsize_ += v.vsize;
    }
  if (n_rec_vars == 1)
    for (Variable& v : vars_)
      if (v.is_record) recsize_ = fixed_items(v) * type_size(v.type);
  Out hdr;
  build(hdr);
  write_at(0, hdr.buf.data(), hdr.buf.size());
  // zero-fill the fixed part so that the file is valid even before every variable is written
  std::vector<char> zeros(1 << 20, 0);
  int64_t           pos = (int64_t)hdr.buf.size();
  while (pos < rec_begin_) {
    size_t n = (size_t)std::min<int64_t>((int64_t)zeros.size(), rec_begin_ - pos);
    write_at(pos, zeros.data(), n);
    pos += (int64_t)n;
  }
}

void
Writer::put_double(const std::string& var, const double* data, int64_t n)
{
  end_define();
  Variable& v = find(var);
  if (v.is_record || v.type != DOUBLE || n != fixed_items(v)) throw std::runtime_error("netcdf3: put_double mismatch on " + var);
  std::vector<char> tmp((size_t)n * 8);
  memcpy(tmp.data(), data, tmp.size());
  swap_bytes(tmp.data(), 8, n);
  write_at(v.begin, tmp.data(), tmp.size());
}

void
Writer::put_int(const std::string& var, const int* data, int64_t n)
{
  end_define();
  Variable& v = find(var);
  if (v.is_record || v.type != INT || n != fixed_items(v)) throw std::runtime_error("netcdf3: put_int mismatch on " + var);
  std::vector<char> tmp((size_t)n * 4);
  memcpy(tmp.data(), data, tmp.size());
  swap_bytes(tmp.data(), 4, n);
  write_at(v.begin, tmp.data(), tmp.size());
}

void
Writer::put_strings(const std::string& var, const std::vector<std::string>& rows)
{
  end_define();
  Variable& v = find(var);
  if (v.is_record || v.type != CHAR || v.dim_ids.empty()) throw std::runtime_error("netcdf3: put_strings mismatch on " + var);
  const int64_t len   = dim_lens_[v.dim_ids.back()];
  const int64_t nrows = fixed_items(v) / (len ? len : 1);
  if ((int64_t)rows.size() > nrows) throw std::runtime_error("netcdf3: too many rows for " + var);
  std::vector<char> tmp((size_t)(nrows * len), 0);
  for (size_t r = 0; r < rows.size(); ++r) memcpy(tmp.data() + r * len, rows[r].data(), std::min<size_t>(rows[r].size(), (size_t)len - 1));
  write_at(v.begin, tmp.data(), tmp.size());
}

void
Writer::ensure_records(int64_t n)
{
  if (n <= numrecs_) return;
  // extend the file to the end of the new records (the file system supplies the zeros: one byte at the end instead
  // of writing every record twice), then publish the record count (offset 4 of the header)
  if (recsize_ > 0) {
    const char zero = 0;
    write_at(rec_begin_ + n * recsize_ - 1, &zero, 1);
  }
  numrecs_ = n;
  Out o;
  o.i32((int32_t)numrecs_);
  write_at(4, o.buf.data(), 4);
}

void
Writer::put_record_double(const std::string& var, int64_t record, const double* data, int64_t n)
{
  end_define();
  Variable& v = find(var);
  if (!v.is_record || v.type != DOUBLE || n != fixed_items(v))
    throw std::runtime_error("netcdf3: put_record_double mismatch on " + var);
  ensure_records(record + 1);
  std::vector<char> tmp((size_t)n * 8);
  memcpy(tmp.data(), data, tmp.size());
  swap_bytes(tmp.data(), 8, n);
  write_at(v.begin + record * recsize_, tmp.data(), tmp.size());
}

}  // namespace nc3
}  // namespace nimble_b200
