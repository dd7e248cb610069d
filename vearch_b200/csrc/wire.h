// Wire formats of the gamma C-ABI payloads, hand-written because neither libprotobuf nor
// flatc/flatbuffers are available in this build:
//   * flatbuffers reader + builder for gamma_api.Table / gamma_api.Doc
//     (internal/engine/idl/fbs/table.fbs, doc.fbs; consumed by c_api/api_data/table.cc:30-157 and
//     doc.cc:16-76; produced on the Go side by sdk/go/gamma/{table,doc}.go)
//   * proto3 wire codec for vearchpb.SearchRequest (in) / SearchResponse (out)
//     (internal/proto/router_grpc.proto:168-219, data_model.proto:32-37; consumed by
//     c_api/api_data/request.cc:17-91, produced by response.cc:46-185)
// Readers are bounds-checked: a malformed buffer yields "absent", never an out-of-range read.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace gb {

// ------------------------------------------------------------------------------------------
// flatbuffers
// ------------------------------------------------------------------------------------------
class FbTable {
 public:
  FbTable() {}
  FbTable(const uint8_t* buf, size_t len, size_t pos) : buf_(buf), len_(len), pos_(pos) {
    int32_t so;
    if (!rd(pos_, &so)) {
      ok_ = false;
      return;
    }
    int64_t vt = (int64_t)pos_ - so;
    uint16_t vtsize;
    if (vt < 0 || !rd((size_t)vt, &vtsize) || vtsize < 4 || (size_t)vt + vtsize > len_) {
      ok_ = false;
      return;
    }
    vt_ = (size_t)vt;
    vtsize_ = vtsize;
    ok_ = true;
  }
  static FbTable root(const uint8_t* buf, size_t len) {
    uint32_t off;
    if (len < 8) return FbTable();
    memcpy(&off, buf, 4);
    if (off >= len) return FbTable();
    return FbTable(buf, len, off);
  }
  bool ok() const { return ok_; }

  template <typename T>
  T scalar(int id, T def) const {
    size_t o = field_off(id);
    T v;
    if (!o || !rd(pos_ + o, &v)) return def;
    return v;
  }
  // string or [ubyte]: both are u32 length + bytes
  bool bytes(int id, const uint8_t** p, size_t* n) const {
    size_t t = indirect(id);
    uint32_t l;
    if (!t || !rd(t, &l) || t + 4 + (size_t)l > len_) return false;
    *p = buf_ + t + 4;
    *n = l;
    return true;
  }
  std::string str(int id) const {
    const uint8_t* p;
    size_t n;
    if (!bytes(id, &p, &n)) return std::string();
    return std::string(reinterpret_cast<const char*>(p), n);
  }
  bool has(int id) const { return field_off(id) != 0; }
  // vector of tables
  size_t vec_len(int id) const {
    size_t t = indirect(id);
    uint32_t l;
    if (!t || !rd(t, &l) || t + 4 + (size_t)l * 4 > len_) return 0;
    return l;
  }
  FbTable vec_table(int id, size_t i) const {
    size_t t = indirect(id);
    if (!t) return FbTable();
    size_t e = t + 4 + i * 4;
    uint32_t off;
    if (!rd(e, &off) || e + off >= len_) return FbTable();
    return FbTable(buf_, len_, e + off);
  }
  std::string vec_str(int id, size_t i) const {
    size_t t = indirect(id);
    if (!t) return std::string();
    size_t e = t + 4 + i * 4;
    uint32_t off, l;
    if (!rd(e, &off) || !rd(e + off, &l) || e + off + 4 + (size_t)l > len_) return std::string();
    return std::string(reinterpret_cast<const char*>(buf_ + e + off + 4), l);
  }

 private:
  template <typename T>
  bool rd(size_t at, T* v) const {
    if (at + sizeof(T) > len_) return false;
    memcpy(v, buf_ + at, sizeof(T));
    return true;
  }
  size_t field_off(int id) const {
    if (!ok_) return 0;
    size_t slot = 4 + 2 * (size_t)id;
    if (slot + 2 > vtsize_) return 0;
    uint16_t o;
    if (!rd(vt_ + slot, &o)) return 0;
    return o;
  }
  size_t indirect(int id) const {
    size_t o = field_off(id);
    uint32_t u;
    if (!o || !rd(pos_ + o, &u)) return 0;
    size_t t = pos_ + o + u;
    return t < len_ ? t : 0;
  }
  const uint8_t* buf_ = nullptr;
  size_t len_ = 0, pos_ = 0, vt_ = 0, vtsize_ = 0;
  bool ok_ = false;
};

// Minimal back-to-front flatbuffers builder (same layout rules as the official builders, so
// the Go/C++/Python readers accept its output).
class FbBuilder {
 public:
  typedef uint32_t Off;  // offset from the END of the buffer
  Off create_bytes(const void* data, size_t len, bool is_string) {
    prealign(len + (is_string ? 1 : 0), 4);
    if (is_string) push_byte(0);
    push_raw(data, len);
    push<uint32_t>((uint32_t)len);
    return (Off)used_;
  }
  Off create_string(const std::string& s) { return create_bytes(s.data(), s.size(), true); }
  Off create_offset_vector(const std::vector<Off>& offs) {
    prealign(offs.size() * 4, 4);
    for (size_t i = offs.size(); i-- > 0;) push<uint32_t>(refer_to(offs[i]));
    push<uint32_t>((uint32_t)offs.size());
    return (Off)used_;
  }
  void start_table(int nfields) {
    field_loc_.assign(nfields, 0);
    object_start_ = used_;
  }
  void add_offset(int id, Off off) {
    if (!off) return;
    push<uint32_t>(refer_to(off));
    field_loc_[id] = used_;
  }
  template <typename T>
  void add_scalar(int id, T v) {
    push<T>(v);
    field_loc_[id] = used_;
  }
  Off end_table() {
    push<int32_t>(0);  // soffset to the vtable, patched below
    size_t table_start = used_;
    size_t nf = field_loc_.size();
    while (nf > 0 && field_loc_[nf - 1] == 0) nf--;
    for (size_t i = nf; i-- > 0;)
      push<uint16_t>(field_loc_[i] ? (uint16_t)(table_start - field_loc_[i]) : (uint16_t)0);
    push<uint16_t>((uint16_t)(table_start - object_start_));
    push<uint16_t>((uint16_t)((nf + 2) * 2));
    int32_t so = (int32_t)(used_ - table_start);
    memcpy(&buf_[buf_.size() - table_start], &so, 4);
    return (Off)table_start;
  }
  void finish(Off root) {
    prealign(4, minalign_);
    push<uint32_t>(refer_to(root));
  }
  const uint8_t* data() const { return buf_.data() + buf_.size() - used_; }
  size_t size() const { return used_; }

 private:
  void reserve(size_t n) {
    if (used_ + n <= buf_.size()) return;
    size_t ns = buf_.size() ? buf_.size() : 256;
    while (ns < used_ + n) ns *= 2;
    std::vector<uint8_t> nb(ns, 0);
    if (used_) memcpy(nb.data() + ns - used_, buf_.data() + buf_.size() - used_, used_);
    buf_.swap(nb);
  }
  void push_byte(uint8_t b) {
    reserve(1);
    used_++;
    buf_[buf_.size() - used_] = b;
  }
  void push_raw(const void* p, size_t n) {
    reserve(n);
    used_ += n;
    if (n) memcpy(&buf_[buf_.size() - used_], p, n);
  }
  void pad(size_t n) {
    for (size_t i = 0; i < n; i++) push_byte(0);
  }
  void align(size_t a) {
    if (a > minalign_) minalign_ = a;
    pad((~used_ + 1) & (a - 1));
  }
  void prealign(size_t len, size_t a) {
    if (a > minalign_) minalign_ = a;
    pad((~(used_ + len) + 1) & (a - 1));
  }
  template <typename T>
  void push(T v) {
    align(sizeof(T));
    push_raw(&v, sizeof(T));
  }
  uint32_t refer_to(Off off) {
    align(4);
    return (uint32_t)(used_ - off + 4);
  }
  std::vector<uint8_t> buf_;
  size_t used_ = 0, minalign_ = 1, object_start_ = 0;
  std::vector<size_t> field_loc_;
};

// ------------------------------------------------------------------------------------------
// protobuf (proto3 wire format)
// ------------------------------------------------------------------------------------------
struct PbField {
  uint32_t num = 0;
  int wire = 0;  // 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32
  uint64_t val = 0;
  const uint8_t* data = nullptr;
  size_t len = 0;
};

class PbReader {
 public:
  PbReader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
  // returns false at end of message or on malformed input (check error())
  bool next(PbField* f) {
    if (p_ >= end_) return false;
    uint64_t key;
    if (!varint(&key)) return fail();
    f->num = (uint32_t)(key >> 3);
    f->wire = (int)(key & 7);
    f->data = nullptr;
    f->len = 0;
    f->val = 0;
    switch (f->wire) {
      case 0: return varint(&f->val) ? true : fail();
      case 1:
        if (end_ - p_ < 8) return fail();
        memcpy(&f->val, p_, 8);
        p_ += 8;
        return true;
      case 5: {
        if (end_ - p_ < 4) return fail();
        uint32_t v;
        memcpy(&v, p_, 4);
        f->val = v;
        p_ += 4;
        return true;
      }
      case 2: {
        uint64_t l;
        if (!varint(&l) || l > (uint64_t)(end_ - p_)) return fail();
        f->data = p_;
        f->len = (size_t)l;
        p_ += l;
        return true;
      }
      default: return fail();
    }
  }
  bool error() const { return err_; }

 private:
  bool fail() {
    err_ = true;
    p_ = end_;
    return false;
  }
  bool varint(uint64_t* v) {
    uint64_t r = 0;
    for (int shift = 0; shift < 64 && p_ < end_; shift += 7) {
      uint8_t b = *p_++;
      r |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) {
        *v = r;
        return true;
      }
    }
    return false;
  }
  const uint8_t* p_;
  const uint8_t* end_;
  bool err_ = false;
};

class PbWriter {
 public:
  std::string out;
  void varint(uint64_t v) {
    while (v >= 0x80) {
      out.push_back((char)((v & 0x7F) | 0x80));
      v >>= 7;
    }
    out.push_back((char)v);
  }
  void key(uint32_t num, int wire) { varint(((uint64_t)num << 3) | (uint64_t)wire); }
  // proto3: scalar fields equal to their default are not emitted (matches golang/protobuf and
  // the C++ library, so the bytes are identical to what the reference engine produces)
  void put_varint(uint32_t num, uint64_t v) {
    if (!v) return;
    key(num, 0);
    varint(v);
  }
  void put_int32(uint32_t num, int32_t v) { put_varint(num, (uint64_t)(int64_t)v); }
  void put_bool(uint32_t num, bool v) { put_varint(num, v ? 1 : 0); }
  void put_double(uint32_t num, double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    if (!b) return;
    key(num, 1);
    out.append(reinterpret_cast<const char*>(&b), 8);
  }
  void put_bytes(uint32_t num, const void* p, size_t n, bool emit_empty = false) {
    if (!n && !emit_empty) return;
    key(num, 2);
    varint(n);
    out.append(reinterpret_cast<const char*>(p), n);
  }
  void put_string(uint32_t num, const std::string& s) { put_bytes(num, s.data(), s.size()); }
  void put_message(uint32_t num, const std::string& body) { put_bytes(num, body.data(), body.size(), true); }
};

inline double pb_double(uint64_t bits) {
  double d;
  memcpy(&d, &bits, 8);
  return d;
}

}  // namespace gb
