// proto_lite: a protobuf-free runtime for the subset of caffe.proto the DeeperCut path uses.
//
// The reference links libprotobuf + protoc-generated caffe.pb.{h,cc} (src/caffe/proto/caffe.proto,
// src/caffe/util/io.cpp:34-65); neither exists in this image.  Messages here are plain structs
// declared with X-macro field lists (caffe.pb.h) that expand to protobuf-style accessors
// (name(), has_name(), set_name(), foo_size(), foo(i), add_foo(), mutable_foo()), a text-format
// parser (TextFormat::Parse subset: comments, '/" strings, {} and <> nesting, repeated scalars,
// [a, b] lists, enum identifiers) and a binary wire-format reader/writer (varint / fixed32 /
// fixed64 / length-delimited incl. packed repeated scalars) so .caffemodel files round-trip.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace caffe {
namespace pl {

// ----------------------------------------------------------------------------- text lexer
struct Token {
  enum Kind { END, IDENT, STRING, PUNCT } kind = END;
  std::string text;
  int line = 0;
};

class TextLexer {
 public:
  explicit TextLexer(const std::string& s) : s_(s) {}
  Token Next();
  const Token& Peek() {
    if (!has_peek_) { peek_ = Next(); has_peek_ = true; }
    return peek_;
  }
  Token Take() {
    if (has_peek_) { has_peek_ = false; return peek_; }
    return Next();
  }
  std::string error;

 private:
  const std::string& s_;
  size_t i_ = 0;
  int line_ = 1;
  bool has_peek_ = false;
  Token peek_;
};

// ----------------------------------------------------------------------------- wire reader
class WireReader {
 public:
  WireReader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
  bool done() const { return p_ >= end_; }
  bool ok() const { return ok_; }
  uint64_t Varint();
  uint32_t Fixed32();
  uint64_t Fixed64();
  WireReader Sub();   // length-delimited payload
  std::string Bytes();
  void Skip(int wire_type);
  size_t remaining() const { return static_cast<size_t>(end_ - p_); }

 private:
  const uint8_t* p_;
  const uint8_t* end_;
  bool ok_ = true;
};

class WireWriter {
 public:
  std::string out;
  void Varint(uint64_t v);
  void Tag(int field, int wt) { Varint(static_cast<uint64_t>(field) << 3 | static_cast<uint64_t>(wt)); }
  void Fixed32(uint32_t v) { out.append(reinterpret_cast<const char*>(&v), 4); }
  void Fixed64(uint64_t v) { out.append(reinterpret_cast<const char*>(&v), 8); }
  void Bytes(int field, const std::string& b) { Tag(field, 2); Varint(b.size()); out += b; }
};

// ----------------------------------------------------------------------------- scalar traits
typedef std::map<std::string, int> EnumTable;

template <class T> struct Scalar;
#define PL_INT_SCALAR(T)                                                                     \
  template <> struct Scalar<T> {                                                             \
    static const int kWire = 0;                                                              \
    static bool FromText(const Token& t, T* v, const EnumTable*);                            \
    static void Read(WireReader& r, int wt, T* v) { (void)wt; *v = static_cast<T>(r.Varint()); } \
    static void Write(WireWriter& w, const T& v) { w.Varint(static_cast<uint64_t>(static_cast<int64_t>(v))); } \
    static std::string ToText(const T& v, const EnumTable*);                                 \
  };
PL_INT_SCALAR(int32_t)
PL_INT_SCALAR(uint32_t)
PL_INT_SCALAR(int64_t)
PL_INT_SCALAR(bool)
#undef PL_INT_SCALAR
template <> struct Scalar<float> {
  static const int kWire = 5;
  static bool FromText(const Token& t, float* v, const EnumTable*);
  static void Read(WireReader& r, int, float* v) { uint32_t b = r.Fixed32(); memcpy(v, &b, 4); }
  static void Write(WireWriter& w, const float& v) { uint32_t b; memcpy(&b, &v, 4); w.Fixed32(b); }
  static std::string ToText(const float& v, const EnumTable*);
};
template <> struct Scalar<double> {
  static const int kWire = 1;
  static bool FromText(const Token& t, double* v, const EnumTable*);
  static void Read(WireReader& r, int, double* v) { uint64_t b = r.Fixed64(); memcpy(v, &b, 8); }
  static void Write(WireWriter& w, const double& v) { uint64_t b; memcpy(&b, &v, 8); w.Fixed64(b); }
  static std::string ToText(const double& v, const EnumTable*);
};
template <> struct Scalar<std::string> {
  static const int kWire = 2;
  static bool FromText(const Token& t, std::string* v, const EnumTable*) { *v = t.text; return t.kind == Token::STRING || t.kind == Token::IDENT; }
  static void Read(WireReader& r, int, std::string* v) { *v = r.Bytes(); }
  static void Write(WireWriter& w, const std::string& v) { w.Varint(v.size()); w.out += v; }
  static std::string ToText(const std::string& v, const EnumTable*);
};

// ----------------------------------------------------------------------------- message base
class Message {
 public:
  virtual ~Message() {}
  // Parses fields until `closer` ('}' / '>' / 0 for top level).  Unknown fields are an error,
  // as with TextFormat::Parse.
  bool ParseText(TextLexer& lx, char closer);
  bool ParseFromTextString(const std::string& s, std::string* err = nullptr);
  bool ParseWire(WireReader& r);
  bool ParseFromBinaryString(const std::string& s) {
    WireReader r(reinterpret_cast<const uint8_t*>(s.data()), s.size());
    return ParseWire(r) && r.ok();
  }
  void SerializeTo(WireWriter& w) const { WriteFields(w); }
  std::string SerializeAsString() const { WireWriter w; WriteFields(w); return w.out; }
  std::string DebugString(int indent = 0) const { std::string s; PrintFields(&s, indent); return s; }
  virtual const char* TypeName() const = 0;

 protected:
  // One text field `name` whose value starts at lx: 0 = parsed, 1 = not a field of this
  // message (the caller skips its value: this schema is a subset of caffe.proto), 2 = bad value.
  virtual int TextField(const std::string& name, TextLexer& lx) = 0;
  virtual bool WireField(int number, int wire_type, WireReader& r) = 0;
  virtual void WriteFields(WireWriter& w) const = 0;
  virtual void PrintFields(std::string* s, int indent) const = 0;
};

// helpers used by the generated code -------------------------------------------------------
template <class T>
bool TextScalar(TextLexer& lx, T* v, const EnumTable* e) {
  Token t = lx.Take();
  return Scalar<T>::FromText(t, v, e);
}
template <class T>
bool TextRepeated(TextLexer& lx, std::vector<T>* v, const EnumTable* e) {
  if (lx.Peek().kind == Token::PUNCT && lx.Peek().text == "[") {
    lx.Take();
    while (true) {
      if (lx.Peek().kind == Token::PUNCT && lx.Peek().text == "]") { lx.Take(); return true; }
      if (lx.Peek().kind == Token::PUNCT && lx.Peek().text == ",") { lx.Take(); continue; }
      if (lx.Peek().kind == Token::END) return false;
      T x{};
      if (!TextScalar(lx, &x, e)) return false;
      v->push_back(x);
    }
  }
  T x{};
  if (!TextScalar(lx, &x, e)) return false;
  v->push_back(x);
  return true;
}
bool TextOpenMessage(TextLexer& lx, char* closer);   // consumes '{' or '<'
template <class T>
void WireRepeated(WireReader& r, int wt, std::vector<T>* v) {
  if (wt == 2 && Scalar<T>::kWire != 2) {            // packed
    WireReader sub = r.Sub();
    while (!sub.done() && sub.ok()) { T x{}; Scalar<T>::Read(sub, Scalar<T>::kWire, &x); v->push_back(x); }
  } else {
    T x{};
    Scalar<T>::Read(r, wt, &x);
    v->push_back(x);
  }
}
void Indent(std::string* s, int n);

}  // namespace pl
}  // namespace caffe

// =============================================================================================
// X-macro expansion.  A message is declared as
//   #define FOO_FIELDS(OPT, REP, MSG, RMSG, ENM)  OPT(float, eps, 3, 1e-5f) REP(uint32_t, pad, 3) ...
//   PL_DECLARE_MESSAGE(Foo, FOO_FIELDS)
// OPT(type, name, number, default)   optional scalar     name() has_name() set_name() clear_name()
// REP(type, name, number, packed)    repeated scalar     name_size() name(i) add_name(v) name() mutable_name()
// MSG(Type, name, number)            optional message    name() has_name() mutable_name() clear_name()
// RMSG(Type, name, number)           repeated message    name_size() name(i) add_name() mutable_name(i)
// ENM(EnumType, name, number, default, table_fn)  optional enum stored as int
// =============================================================================================
#define PL_M_OPT(T, n, num, def)                     \
 private:                                            \
  T n##_ = def;                                      \
  bool has_##n##_ = false;                           \
                                                     \
 public:                                             \
  const T& n() const { return n##_; }                \
  bool has_##n() const { return has_##n##_; }        \
  void set_##n(const T& v) { n##_ = v; has_##n##_ = true; } \
  void clear_##n() { n##_ = def; has_##n##_ = false; }
#define PL_M_REP(T, n, num, packed)                  \
 private:                                            \
  std::vector<T> n##_;                               \
                                                     \
 public:                                             \
  int n##_size() const { return static_cast<int>(n##_.size()); } \
  const T& n(int i) const { return n##_[i]; }        \
  void add_##n(const T& v) { n##_.push_back(v); }    \
  void set_##n(int i, const T& v) { n##_[i] = v; }   \
  const std::vector<T>& n() const { return n##_; }   \
  std::vector<T>* mutable_##n() { return &n##_; }    \
  void clear_##n() { n##_.clear(); }
#define PL_M_MSG(T, n, num)                          \
 private:                                            \
  T n##_;                                            \
  bool has_##n##_ = false;                           \
                                                     \
 public:                                             \
  const T& n() const { return n##_; }                \
  bool has_##n() const { return has_##n##_; }        \
  T* mutable_##n() { has_##n##_ = true; return &n##_; } \
  void clear_##n() { n##_ = T(); has_##n##_ = false; }
#define PL_M_RMSG(T, n, num)                         \
 private:                                            \
  std::vector<T> n##_;                               \
                                                     \
 public:                                             \
  int n##_size() const { return static_cast<int>(n##_.size()); } \
  const T& n(int i) const { return n##_[i]; }        \
  T* mutable_##n(int i) { return &n##_[i]; }         \
  T* add_##n() { n##_.emplace_back(); return &n##_.back(); } \
  const std::vector<T>& n() const { return n##_; }   \
  std::vector<T>* mutable_##n() { return &n##_; }    \
  void clear_##n() { n##_.clear(); }
#define PL_M_ENM(E, n, num, def, tbl)                \
 private:                                            \
  E n##_ = def;                                      \
  bool has_##n##_ = false;                           \
                                                     \
 public:                                             \
  E n() const { return n##_; }                       \
  bool has_##n() const { return has_##n##_; }        \
  void set_##n(E v) { n##_ = v; has_##n##_ = true; } \
  void clear_##n() { n##_ = def; has_##n##_ = false; }

// text
#define PL_T_OPT(T, n, num, def) \
  if (name == #n) { has_##n##_ = true; return ::caffe::pl::TextScalar(lx, &n##_, nullptr) ? 0 : 2; }
#define PL_T_REP(T, n, num, packed) \
  if (name == #n) return ::caffe::pl::TextRepeated(lx, &n##_, nullptr) ? 0 : 2;
#define PL_T_MSG(T, n, num)                                              \
  if (name == #n) {                                                      \
    char c;                                                              \
    if (!::caffe::pl::TextOpenMessage(lx, &c)) return 2;                 \
    has_##n##_ = true;                                                   \
    return n##_.ParseText(lx, c) ? 0 : 2;                                \
  }
#define PL_T_RMSG(T, n, num)                                             \
  if (name == #n) {                                                      \
    char c;                                                              \
    if (!::caffe::pl::TextOpenMessage(lx, &c)) return 2;                 \
    n##_.emplace_back();                                                 \
    return n##_.back().ParseText(lx, c) ? 0 : 2;                         \
  }
#define PL_T_ENM(E, n, num, def, tbl)                                    \
  if (name == #n) {                                                      \
    int32_t v = 0;                                                       \
    if (!::caffe::pl::TextScalar(lx, &v, &tbl())) return 2;              \
    n##_ = static_cast<E>(v);                                            \
    has_##n##_ = true;                                                   \
    return 0;                                                            \
  }
// wire read
#define PL_W_OPT(T, n, num, def) \
  case num: ::caffe::pl::Scalar<T>::Read(r, wt, &n##_); has_##n##_ = true; return true;
#define PL_W_REP(T, n, num, packed) \
  case num: ::caffe::pl::WireRepeated(r, wt, &n##_); return true;
#define PL_W_MSG(T, n, num) \
  case num: { ::caffe::pl::WireReader sub = r.Sub(); has_##n##_ = true; return n##_.ParseWire(sub); }
#define PL_W_RMSG(T, n, num) \
  case num: { ::caffe::pl::WireReader sub = r.Sub(); n##_.emplace_back(); return n##_.back().ParseWire(sub); }
#define PL_W_ENM(E, n, num, def, tbl) \
  case num: n##_ = static_cast<E>(r.Varint()); has_##n##_ = true; return true;
// wire write
#define PL_S_OPT(T, n, num, def) \
  if (has_##n##_) { w.Tag(num, ::caffe::pl::Scalar<T>::kWire); ::caffe::pl::Scalar<T>::Write(w, n##_); }
#define PL_S_REP(T, n, num, packed)                                                        \
  if (!n##_.empty()) {                                                                     \
    if (packed && ::caffe::pl::Scalar<T>::kWire != 2) {                                    \
      ::caffe::pl::WireWriter sub;                                                         \
      for (const auto& v : n##_) ::caffe::pl::Scalar<T>::Write(sub, v);                    \
      w.Bytes(num, sub.out);                                                               \
    } else {                                                                               \
      for (const auto& v : n##_) { w.Tag(num, ::caffe::pl::Scalar<T>::kWire); ::caffe::pl::Scalar<T>::Write(w, v); } \
    }                                                                                      \
  }
#define PL_S_MSG(T, n, num) \
  if (has_##n##_) w.Bytes(num, n##_.SerializeAsString());
#define PL_S_RMSG(T, n, num) \
  for (const auto& m : n##_) w.Bytes(num, m.SerializeAsString());
#define PL_S_ENM(E, n, num, def, tbl) \
  if (has_##n##_) { w.Tag(num, 0); w.Varint(static_cast<uint64_t>(static_cast<int64_t>(n##_))); }
// print
#define PL_P_OPT(T, n, num, def) \
  if (has_##n##_) { ::caffe::pl::Indent(s, indent); *s += #n ": " + ::caffe::pl::Scalar<T>::ToText(n##_, nullptr) + "\n"; }
#define PL_P_REP(T, n, num, packed) \
  for (const auto& v : n##_) { ::caffe::pl::Indent(s, indent); *s += #n ": " + ::caffe::pl::Scalar<T>::ToText(v, nullptr) + "\n"; }
#define PL_P_MSG(T, n, num)                                                   \
  if (has_##n##_) {                                                           \
    ::caffe::pl::Indent(s, indent); *s += #n " {\n";                          \
    *s += n##_.DebugString(indent + 2);                                       \
    ::caffe::pl::Indent(s, indent); *s += "}\n";                              \
  }
#define PL_P_RMSG(T, n, num)                                                  \
  for (const auto& m : n##_) {                                                \
    ::caffe::pl::Indent(s, indent); *s += #n " {\n";                          \
    *s += m.DebugString(indent + 2);                                          \
    ::caffe::pl::Indent(s, indent); *s += "}\n";                              \
  }
#define PL_P_ENM(E, n, num, def, tbl) \
  if (has_##n##_) { ::caffe::pl::Indent(s, indent); *s += #n ": " + ::caffe::pl::Scalar<int32_t>::ToText(static_cast<int32_t>(n##_), &tbl()) + "\n"; }

#define PL_DECLARE_MESSAGE(Name, FIELDS)                                                   \
  class Name : public ::caffe::pl::Message {                                               \
    FIELDS(PL_M_OPT, PL_M_REP, PL_M_MSG, PL_M_RMSG, PL_M_ENM)                              \
   public:                                                                                 \
    const char* TypeName() const override { return #Name; }                                \
    void CopyFrom(const Name& o) { *this = o; }                                            \
    void Clear() { *this = Name(); }                                                       \
                                                                                           \
   protected:                                                                              \
    int TextField(const std::string& name, ::caffe::pl::TextLexer& lx) override {          \
      FIELDS(PL_T_OPT, PL_T_REP, PL_T_MSG, PL_T_RMSG, PL_T_ENM)                            \
      return 1;                                                                            \
    }                                                                                      \
    bool WireField(int number, int wt, ::caffe::pl::WireReader& r) override {              \
      switch (number) {                                                                    \
        FIELDS(PL_W_OPT, PL_W_REP, PL_W_MSG, PL_W_RMSG, PL_W_ENM)                          \
        default: break;                                                                    \
      }                                                                                    \
      return false;                                                                        \
    }                                                                                      \
    void WriteFields(::caffe::pl::WireWriter& w) const override {                          \
      FIELDS(PL_S_OPT, PL_S_REP, PL_S_MSG, PL_S_RMSG, PL_S_ENM)                            \
    }                                                                                      \
    void PrintFields(std::string* s, int indent) const override {                          \
      FIELDS(PL_P_OPT, PL_P_REP, PL_P_MSG, PL_P_RMSG, PL_P_ENM)                            \
    }                                                                                      \
  };
