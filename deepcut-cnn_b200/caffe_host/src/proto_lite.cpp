// Runtime of proto_lite: text lexer/parser, wire reader/writer, enum tables.
#include "caffe/proto/caffe.pb.h"

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace caffe {
namespace pl {

static bool IsPunct(char c) { return strchr("{}<>[]:,;", c) != nullptr; }

Token TextLexer::Next() {
  const size_t n = s_.size();
  while (i_ < n) {
    const char c = s_[i_];
    if (c == '#') { while (i_ < n && s_[i_] != '\n') ++i_; }
    else if (c == '\n') { ++line_; ++i_; }
    else if (c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v') ++i_;
    else break;
  }
  Token t;
  t.line = line_;
  if (i_ >= n) return t;
  const char c = s_[i_];
  if (IsPunct(c)) { t.kind = Token::PUNCT; t.text.assign(1, c); ++i_; return t; }
  if (c == '"' || c == '\'') {
    t.kind = Token::STRING;
    size_t j = i_ + 1;
    while (j < n && s_[j] != c) {
      if (s_[j] == '\\' && j + 1 < n) {
        ++j;
        switch (s_[j]) { case 'n': t.text += '\n'; break; case 't': t.text += '\t'; break; default: t.text += s_[j]; }
      } else {
        if (s_[j] == '\n') ++line_;
        t.text += s_[j];
      }
      ++j;
    }
    if (j >= n) { error = "unterminated string at line " + std::to_string(t.line); t.kind = Token::END; i_ = n; return t; }
    i_ = j + 1;
    return t;
  }
  t.kind = Token::IDENT;
  size_t j = i_;
  while (j < n && !isspace(static_cast<unsigned char>(s_[j])) && !IsPunct(s_[j]) && s_[j] != '"' && s_[j] != '\'' && s_[j] != '#') ++j;
  t.text = s_.substr(i_, j - i_);
  i_ = j;
  return t;
}

// ----------------------------------------------------------------------------- wire
uint64_t WireReader::Varint() {
  uint64_t v = 0;
  for (int shift = 0; shift < 64; shift += 7) {
    if (p_ >= end_) { ok_ = false; return 0; }
    const uint8_t b = *p_++;
    v |= static_cast<uint64_t>(b & 0x7F) << shift;
    if (!(b & 0x80)) return v;
  }
  ok_ = false;
  return 0;
}
uint32_t WireReader::Fixed32() {
  if (remaining() < 4) { ok_ = false; p_ = end_; return 0; }
  uint32_t v; memcpy(&v, p_, 4); p_ += 4; return v;
}
uint64_t WireReader::Fixed64() {
  if (remaining() < 8) { ok_ = false; p_ = end_; return 0; }
  uint64_t v; memcpy(&v, p_, 8); p_ += 8; return v;
}
WireReader WireReader::Sub() {
  const uint64_t len = Varint();
  if (!ok_ || len > remaining()) { ok_ = false; p_ = end_; return WireReader(end_, 0); }
  WireReader sub(p_, static_cast<size_t>(len));
  p_ += len;
  return sub;
}
std::string WireReader::Bytes() {
  WireReader sub = Sub();
  return std::string(reinterpret_cast<const char*>(sub.p_), sub.remaining());
}
void WireReader::Skip(int wt) {
  switch (wt) {
    case 0: Varint(); break;
    case 1: Fixed64(); break;
    case 2: Sub(); break;
    case 5: Fixed32(); break;
    default: ok_ = false; p_ = end_;
  }
}
void WireWriter::Varint(uint64_t v) {
  while (v >= 0x80) { out.push_back(static_cast<char>((v & 0x7F) | 0x80)); v >>= 7; }
  out.push_back(static_cast<char>(v));
}

// ----------------------------------------------------------------------------- scalars
static bool ParseInt(const Token& t, long long* out, const EnumTable* e) {
  if (t.kind != Token::IDENT) return false;
  if (e) {
    auto it = e->find(t.text);
    if (it != e->end()) { *out = it->second; return true; }
  }
  if (t.text == "true") { *out = 1; return true; }
  if (t.text == "false") { *out = 0; return true; }
  errno = 0;
  char* end = nullptr;
  const long long v = strtoll(t.text.c_str(), &end, 0);
  if (errno || end == t.text.c_str() || *end) return false;
  *out = v;
  return true;
}
static bool ParseDouble(const Token& t, double* out) {
  if (t.kind != Token::IDENT) return false;
  std::string s = t.text;
  if (!s.empty() && (s.back() == 'f' || s.back() == 'F') && s.find_first_of("xX") == std::string::npos) s.pop_back();
  if (s == "inf" || s == "infinity") { *out = INFINITY; return true; }
  if (s == "-inf" || s == "-infinity") { *out = -INFINITY; return true; }
  if (s == "nan") { *out = NAN; return true; }
  char* end = nullptr;
  const double v = strtod(s.c_str(), &end);
  if (end == s.c_str() || *end) return false;
  *out = v;
  return true;
}
#define PL_INT_IMPL(T)                                                              \
  bool Scalar<T>::FromText(const Token& t, T* v, const EnumTable* e) {              \
    long long x;                                                                    \
    if (!ParseInt(t, &x, e)) return false;                                          \
    *v = static_cast<T>(x);                                                         \
    return true;                                                                    \
  }                                                                                 \
  std::string Scalar<T>::ToText(const T& v, const EnumTable* e) {                   \
    if (e) for (const auto& kv : *e) if (kv.second == static_cast<int>(v)) return kv.first; \
    return std::to_string(static_cast<long long>(v));                               \
  }
PL_INT_IMPL(int32_t)
PL_INT_IMPL(uint32_t)
PL_INT_IMPL(int64_t)
#undef PL_INT_IMPL
bool Scalar<bool>::FromText(const Token& t, bool* v, const EnumTable*) {
  long long x;
  if (!ParseInt(t, &x, nullptr)) return false;
  *v = x != 0;
  return true;
}
std::string Scalar<bool>::ToText(const bool& v, const EnumTable*) { return v ? "true" : "false"; }
bool Scalar<float>::FromText(const Token& t, float* v, const EnumTable*) {
  double d;
  if (!ParseDouble(t, &d)) return false;
  *v = static_cast<float>(d);
  return true;
}
std::string Scalar<float>::ToText(const float& v, const EnumTable*) { char b[40]; snprintf(b, sizeof(b), "%.9g", v); return b; }
bool Scalar<double>::FromText(const Token& t, double* v, const EnumTable*) { return ParseDouble(t, v); }
std::string Scalar<double>::ToText(const double& v, const EnumTable*) { char b[48]; snprintf(b, sizeof(b), "%.17g", v); return b; }
std::string Scalar<std::string>::ToText(const std::string& v, const EnumTable*) {
  std::string o = "\"";
  for (char c : v) { if (c == '"' || c == '\\') o += '\\'; if (c == '\n') { o += "\\n"; continue; } o += c; }
  return o + "\"";
}

void Indent(std::string* s, int n) { s->append(static_cast<size_t>(n), ' '); }

bool TextOpenMessage(TextLexer& lx, char* closer) {
  Token t = lx.Take();
  if (t.kind != Token::PUNCT) return false;
  if (t.text == "{") { *closer = '}'; return true; }
  if (t.text == "<") { *closer = '>'; return true; }
  return false;
}

// Skips the value of a field this schema does not model (scalar, list or nested message).
static bool SkipValue(TextLexer& lx) {
  Token t = lx.Take();
  if (t.kind == Token::PUNCT && (t.text == "{" || t.text == "<")) {
    const std::string close = t.text == "{" ? "}" : ">";
    int depth = 1;
    while (depth > 0) {
      Token u = lx.Take();
      if (u.kind == Token::END) return false;
      if (u.kind == Token::PUNCT && (u.text == "{" || u.text == "<")) ++depth;
      if (u.kind == Token::PUNCT && (u.text == "}" || u.text == ">")) --depth;
    }
    return true;
  }
  if (t.kind == Token::PUNCT && t.text == "[") {
    while (true) {
      Token u = lx.Take();
      if (u.kind == Token::END) return false;
      if (u.kind == Token::PUNCT && u.text == "]") return true;
    }
  }
  return t.kind == Token::IDENT || t.kind == Token::STRING;
}

bool Message::ParseText(TextLexer& lx, char closer) {
  while (true) {
    Token t = lx.Take();
    if (t.kind == Token::END) {
      if (closer) lx.error = std::string("missing '") + closer + "' in " + TypeName();
      return closer == 0 && lx.error.empty();
    }
    if (t.kind == Token::PUNCT) {
      if (closer && t.text[0] == closer) return true;
      if (t.text == "," || t.text == ";") continue;
      lx.error = "line " + std::to_string(t.line) + ": unexpected '" + t.text + "' in " + TypeName();
      return false;
    }
    if (t.kind != Token::IDENT) {
      lx.error = "line " + std::to_string(t.line) + ": expected a field name in " + TypeName();
      return false;
    }
    const std::string name = t.text;
    if (lx.Peek().kind == Token::PUNCT && lx.Peek().text == ":") lx.Take();
    const int rc = TextField(name, lx);
    if (rc == 1) {
      // protobuf's TextFormat (the reference's ReadProtoFromTextFile, io.cpp:34-44) rejects a field the schema does not have: a
      // misspelt parameter must not silently yield a different net.  The only names skipped are the caffe.proto fields this build
      // deliberately does not model -- parameters of layer types outside the forward path (their layers are refused by the registry
      // anyway) and the training-only propagate_down.
      static const char* const kOutOfScope[] = {
          "propagate_down", "transform_param", "loss_param", "accuracy_param", "argmax_param", "concat_param", "contrastive_loss_param",
          "data_param", "dropout_param", "dummy_data_param", "elu_param", "embed_param", "exp_param", "flatten_param", "hdf5_data_param",
          "hdf5_output_param", "hinge_loss_param", "image_data_param", "infogain_loss_param", "inner_product_param", "log_param", "lrn_param",
          "memory_data_param", "mvn_param", "power_param", "prelu_param", "python_param", "reduction_param", "reshape_param", "softmax_param",
          "spp_param", "slice_param", "tanh_param", "threshold_param", "tile_param", "window_data_param", "pose_data_param",
          "softmax_with_loss_vec_param"};
      bool skippable = false;
      if (std::string(TypeName()) == "LayerParameter")
        for (const char* k : kOutOfScope) skippable = skippable || name == k;
      if (!skippable) {
        lx.error = "line " + std::to_string(t.line) + ": message " + TypeName() + " has no field named '" + name + "'";
        return false;
      }
      if (!SkipValue(lx)) { lx.error = "line " + std::to_string(t.line) + ": cannot skip field '" + name + "'"; return false; }
      fprintf(stderr, "proto_lite: LayerParameter.%s (line %d) belongs to a layer type outside the forward path -- skipped\n", name.c_str(), t.line);
    } else if (rc == 2) {
      if (lx.error.empty()) lx.error = "line " + std::to_string(t.line) + ": bad value for " + TypeName() + "." + name;
      return false;
    }
  }
}

bool Message::ParseFromTextString(const std::string& s, std::string* err) {
  TextLexer lx(s);
  const bool ok = ParseText(lx, 0) && lx.error.empty();
  if (!ok && err) *err = lx.error.empty() ? "text-format parse error" : lx.error;
  return ok;
}

bool Message::ParseWire(WireReader& r) {
  while (!r.done() && r.ok()) {
    const uint64_t tag = r.Varint();
    if (!r.ok()) return false;
    const int number = static_cast<int>(tag >> 3), wt = static_cast<int>(tag & 7);
    // WireField returns false for numbers this schema does not model: protobuf ignores them
    size_t before = r.remaining();
    if (!WireField(number, wt, r)) {
      if (r.remaining() == before) r.Skip(wt);
    }
  }
  return r.ok();
}

}  // namespace pl

#define PL_TABLE(fn, ...) \
  const pl::EnumTable& fn() { static const pl::EnumTable t = __VA_ARGS__; return t; }
PL_TABLE(Phase_table, {{"TRAIN", 0}, {"TEST", 1}})
PL_TABLE(VarianceNorm_table, {{"FAN_IN", 0}, {"FAN_OUT", 1}, {"AVERAGE", 2}})
PL_TABLE(DimCheckMode_table, {{"STRICT", 0}, {"PERMISSIVE", 1}})
PL_TABLE(Engine_table, {{"DEFAULT", 0}, {"CAFFE", 1}, {"CUDNN", 2}})
PL_TABLE(EltwiseOp_table, {{"PROD", 0}, {"SUM", 1}, {"MAX", 2}})
PL_TABLE(PoolMethod_table, {{"MAX", 0}, {"AVE", 1}, {"STOCHASTIC", 2}})

}  // namespace caffe
